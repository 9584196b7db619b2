// tokenizer.cpp — the slice of the Hugging Face `tokenizers` crate (0.15.2, Cargo.lock:1410) the
// reference uses: Tokenizer::from_file (/root/reference/src/main.rs:580), token_to_id (:531) and
// decode(ids, skip_special_tokens=true) (:640).  Encoding text is never needed, so only the
// id->token table, the added-token list and the GPT-2 byte-level decoder are restated here.
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

#include "../common.h"
#include "json.h"
#include "utf8.h"

struct wb_tokenizer {
    std::vector<std::string> id_to_token;
    std::vector<unsigned char> is_special, is_added, present;
    std::unordered_map<std::string, int64_t> token_to_id;
    int byte_of_char[512];          // GPT-2 bytes_to_unicode inverse; -1 = not a byte-level char
    bool byte_level = true;         // "decoder": {"type": "ByteLevel"}; false = no decoder section (tokens joined by ' ')
};

namespace {

void build_byte_map(wb_tokenizer& t) {
    // GPT-2 bytes_to_unicode: printable bytes map to themselves, the rest to U+0100.. in order
    for (int& v : t.byte_of_char) v = -1;
    int n = 0;
    for (int b = 0; b < 256; ++b) {
        bool printable = (b >= 0x21 && b <= 0x7E) || (b >= 0xA1 && b <= 0xAC) || (b >= 0xAE && b <= 0xFF);
        if (printable) t.byte_of_char[b] = b;
        else t.byte_of_char[256 + n++] = b;
    }
}

std::string decode_ids(const wb_tokenizer& t, const int64_t* ids, int n, bool skip_special) {
    std::string bytes;
    bool first = true;
    for (int i = 0; i < n; ++i) {
        // main.rs:639: ids that do not fit u32 are dropped; unknown ids are skipped by the crate
        if (ids[i] < 0 || ids[i] >= (int64_t)t.id_to_token.size()) continue;
        const size_t id = (size_t)ids[i];
        if (!t.present[id]) continue;             // a hole in the id space: id_to_token() is None
        if (skip_special && t.is_special[id]) continue;
        const std::string& tok = t.id_to_token[id];
        // every surviving token goes through the ByteLevel decoder, added (non-special) ones included, exactly as
        // tokenizers' decode_chain does: a token whose characters are all byte-level characters becomes those bytes,
        // any other token is taken as its raw UTF-8 text
        if (!t.byte_level) {                     // Tokenizer::decode without a decoder: tokens.join(" ")
            if (!first) bytes += ' ';
            first = false;
            bytes += tok;
            continue;
        }
        std::string piece;
        bool ok = true;
        size_t j = 0;
        while (j < tok.size()) {
            uint32_t cp = wbutf8::decode(tok, j);
            if (cp < 512 && t.byte_of_char[cp] >= 0) piece += (char)t.byte_of_char[cp];
            else { ok = false; break; }
        }
        bytes += ok ? piece : tok;      // ByteLevel decoder falls back to the raw token text
    }
    return wbutf8::from_utf8_lossy(bytes);     // invalid sequences become U+FFFD, as in the tokenizers crate
}

}  // namespace

extern "C" {

int wb_tokenizer_load(wb_tokenizer** out, const char* path) {
    try {
        WB_REQUIRE(out && path, WB_EINVAL, "null argument");
        std::ifstream f(path, std::ios::binary);
        WB_REQUIRE(f.good(), WB_EIO, "Failed to load tokenizer %s: cannot open", path);
        std::stringstream ss;
        ss << f.rdbuf();
        wbjson::Value root;
        try {
            root = wbjson::parse(ss.str());
        } catch (const std::exception& e) {
            WB_THROW(WB_EINVAL, "Failed to load tokenizer %s: %s", path, e.what());
        }
        auto* t = new wb_tokenizer();
        build_byte_map(*t);
        auto put = [&](int64_t id, const std::string& tok, bool special, bool added) {
            if (id < 0 || id > (1 << 24)) return;
            if ((size_t)id >= t->id_to_token.size()) {
                t->id_to_token.resize((size_t)id + 1);
                t->is_special.resize((size_t)id + 1, 0);
                t->is_added.resize((size_t)id + 1, 0);
                t->present.resize((size_t)id + 1, 0);
            }
            t->present[(size_t)id] = 1;
            t->id_to_token[(size_t)id] = tok;
            t->is_special[(size_t)id] = special;
            t->is_added[(size_t)id] = added;
            t->token_to_id[tok] = id;
        };
        const wbjson::Value& vocab = root["model"]["vocab"];
        for (const auto& kv : vocab.o) put((int64_t)kv.second.num(), kv.first, false, false);
        for (const auto& a : root["added_tokens"].arr())
            put((int64_t)a["id"].num(), a["content"].str(), a["special"].type == wbjson::Value::Bool && a["special"].b, true);
        const wbjson::Value& dec = root["decoder"];
        // the reference's model (openai/whisper-*) ships a ByteLevel decoder; a file without a decoder section decodes
        // to the tokens joined by ' ' (tokenizers: Tokenizer::decode); other decoder types are not restated here
        if (dec.is_null()) t->byte_level = false;
        else {
            WB_REQUIRE(dec["type"].str() == "ByteLevel", WB_EINVAL, "Failed to load tokenizer %s: decoder type '%s' is not supported (ByteLevel only)",
                       path, dec["type"].str().c_str());
            t->byte_level = true;
        }
        WB_REQUIRE(!t->id_to_token.empty(), WB_EINVAL, "Failed to load tokenizer %s: empty vocabulary", path);
        *out = t;
        return WB_OK;
    } catch (const WbError& e) {
        wb_set_error(e.what());
        return e.code;
    } catch (const std::exception& e) {
        wb_set_error(e.what());
        return WB_EINVAL;
    }
}

void wb_tokenizer_free(wb_tokenizer* t) { delete t; }

int64_t wb_tokenizer_token_to_id(const wb_tokenizer* t, const char* token) {
    if (!t || !token) return -1;
    auto it = t->token_to_id.find(token);
    return it == t->token_to_id.end() ? -1 : it->second;
}

// main.rs:528-569
int wb_host_special_tokens(const wb_tokenizer* tok, const char* language, const char* task, int64_t* o) {
    try {
        WB_REQUIRE(language && task && o, WB_EINVAL, "null argument");
        if (tok) {
            auto get = [&](const std::string& s) {
                int64_t id = wb_tokenizer_token_to_id(tok, s.c_str());
                WB_REQUIRE(id >= 0, WB_EINVAL, "Tokenizer missing token: %s", s.c_str());
                return id;
            };
            o[0] = get("<|startoftranscript|>");
            o[1] = get("<|endoftext|>");
            o[2] = get(std::string("<|") + language + "|>");
            o[3] = get(std::string("<|") + task + "|>");
            o[4] = get("<|notimestamps|>");
            return WB_OK;
        }
        o[0] = 50258;
        o[1] = 50257;
        o[2] = std::strcmp(language, "hi") == 0 ? 50276 : 50259;            // unknown -> en
        o[3] = std::strcmp(task, "translate") == 0 ? 50358 : 50359;         // unknown -> transcribe
        o[4] = 50363;
        return WB_OK;
    } catch (const WbError& e) {
        wb_set_error(e.what());
        return e.code;
    }
}

// main.rs:637-648
int64_t wb_host_decode_tokens(const wb_tokenizer* tok, const int64_t* tokens, int n, char* out, int64_t cap) {
    std::string s;
    if (tok) {
        s = decode_ids(*tok, tokens, n, true);
    } else {
        s = "[TOKENS:";
        for (int i = 0; i < n && i < 200; ++i) {
            if (i) s += ' ';
            s += std::to_string((long long)tokens[i]);
        }
        s += ']';
    }
    if (out && cap > 0) {
        size_t m = s.size() < (size_t)cap - 1 ? s.size() : (size_t)cap - 1;
        std::memcpy(out, s.data(), m);
        out[m] = '\0';
    }
    return (int64_t)s.size();
}

}  // extern "C"
