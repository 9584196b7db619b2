// onnx.cpp — reads the weights out of the reference's ONNX export without ONNX Runtime or protobuf:
// `<onnx_dir>/encoder_model.onnx` + `decoder_model.onnx`, the files build_session() loads at
// /root/reference/src/main.rs:1099-1108, produced by scripts/export_onnx_whisper.py:20-28
// (optimum `main_export(..., task="automatic-speech-recognition-with-past")`).
//
// Only what the weight path needs is parsed from the protobuf wire format: ModelProto.graph (7),
// GraphProto.node (1) / initializer (5), NodeProto.input (1) / output (2) / op_type (4),
// TensorProto.dims (1) / data_type (2) / float_data (4) / name (8) / raw_data (9) / data_location (14).
//
// Name recovery.  torch.onnx keeps parameter names for tensors an op consumes directly (conv
// weights, biases, LayerNorm, embeddings) but constant-folds every nn.Linear weight into an
// anonymous, TRANSPOSED initializer `onnx::MatMul_<n>` of shape [in, out].  Those are matched to
// modules by graph order: the MatMul nodes appear in the order the HF forward executes them
// (encoder layer: q, k, v, out, fc1, fc2; decoder layer: self q, k, v, out, cross q, k, v, out,
// fc1, fc2; then the tied proj_out).  Shapes are checked for every assignment.
//
// Two more things the real exporter does (found by exporting HF's WhisperForConditionalGeneration with torch.onnx in
// tests/test_onnx_torch_export_cpu.py): identical initializers are de-duplicated into Identity aliases, and the
// encoder's position table is folded into an anonymous `onnx::Add_<n>` initializer.  Both are handled below.
//
// No optimum export of the real checkpoint exists offline: the reader is exercised on files synthesised
// by tests/onnx_writer.py and on torch.onnx exports of a randomly initialised HF model of the toy shape.
#include <cstring>
#include <fstream>
#include <map>
#include <string>
#include <vector>

#include "../common.h"

namespace wbonnx {

struct Tensor {
    std::string name;
    std::vector<int64_t> dims;
    int dtype = 0;                     // 1 f32, 10 f16, 16 bf16
    const unsigned char* raw = nullptr;
    size_t raw_len = 0;
    std::vector<float> float_data;
    int data_location = 0;
    int64_t numel() const { int64_t n = 1; for (auto d : dims) n *= d; return n; }
};
struct Node {
    std::string op;
    std::vector<std::string> in, out;
};
struct Model {
    std::vector<unsigned char> bytes;
    std::map<std::string, Tensor> init;
    std::vector<Node> nodes;
};

struct Reader {
    const unsigned char* p;
    const unsigned char* e;
    bool ok() const { return p < e; }
    uint64_t varint() {
        uint64_t v = 0;
        int sh = 0;
        while (p < e) {
            unsigned char b = *p++;
            v |= (uint64_t)(b & 0x7F) << sh;
            if (!(b & 0x80)) return v;
            sh += 7;
            WB_REQUIRE(sh < 70, WB_EIO, "ONNX: malformed varint");
        }
        WB_THROW(WB_EIO, "ONNX: truncated varint");
    }
    Reader sub() {
        uint64_t n = varint();
        WB_REQUIRE((uint64_t)(e - p) >= n, WB_EIO, "ONNX: truncated length-delimited field");
        Reader r{p, p + n};
        p += n;
        return r;
    }
    void skip(int wt) {
        if (wt == 0) varint();
        else if (wt == 1) { WB_REQUIRE(e - p >= 8, WB_EIO, "ONNX: truncated fixed64"); p += 8; }
        else if (wt == 2) sub();
        else if (wt == 5) { WB_REQUIRE(e - p >= 4, WB_EIO, "ONNX: truncated fixed32"); p += 4; }
        else WB_THROW(WB_EIO, "ONNX: unsupported wire type %d", wt);
    }
};

static Tensor parse_tensor(Reader r) {
    Tensor t;
    while (r.ok()) {
        uint64_t key = r.varint();
        int f = (int)(key >> 3), wt = (int)(key & 7);
        if (f == 1 && wt == 0) t.dims.push_back((int64_t)r.varint());
        else if (f == 1 && wt == 2) { Reader s = r.sub(); while (s.ok()) t.dims.push_back((int64_t)s.varint()); }
        else if (f == 2 && wt == 0) t.dtype = (int)r.varint();
        else if (f == 4 && wt == 2) { Reader s = r.sub(); size_t n = (size_t)(s.e - s.p) / 4; t.float_data.resize(n); std::memcpy(t.float_data.data(), s.p, n * 4); }
        else if (f == 4 && wt == 5) { float v; WB_REQUIRE(r.e - r.p >= 4, WB_EIO, "ONNX: truncated float"); std::memcpy(&v, r.p, 4); r.p += 4; t.float_data.push_back(v); }
        else if (f == 8 && wt == 2) { Reader s = r.sub(); t.name.assign((const char*)s.p, (size_t)(s.e - s.p)); }
        else if (f == 9 && wt == 2) { Reader s = r.sub(); t.raw = s.p; t.raw_len = (size_t)(s.e - s.p); }
        else if (f == 14 && wt == 0) t.data_location = (int)r.varint();
        else r.skip(wt);
    }
    return t;
}

static Node parse_node(Reader r) {
    Node n;
    while (r.ok()) {
        uint64_t key = r.varint();
        int f = (int)(key >> 3), wt = (int)(key & 7);
        if ((f == 1 || f == 2 || f == 4) && wt == 2) {
            Reader s = r.sub();
            std::string v((const char*)s.p, (size_t)(s.e - s.p));
            if (f == 1) n.in.push_back(v); else if (f == 2) n.out.push_back(v); else n.op = v;
        } else r.skip(wt);
    }
    return n;
}

void load(const std::string& path, Model& m) {
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    WB_REQUIRE(f.good(), WB_EIO, "Failed to load %s", path.c_str());
    std::streamsize n = f.tellg();
    f.seekg(0);
    m.bytes.resize((size_t)n);
    f.read((char*)m.bytes.data(), n);
    WB_REQUIRE(f.good(), WB_EIO, "short read: %s", path.c_str());
    Reader r{m.bytes.data(), m.bytes.data() + m.bytes.size()};
    bool have_graph = false;
    while (r.ok()) {
        uint64_t key = r.varint();
        int fld = (int)(key >> 3), wt = (int)(key & 7);
        if (fld == 7 && wt == 2) {                       // ModelProto.graph
            have_graph = true;
            Reader g = r.sub();
            while (g.ok()) {
                uint64_t k2 = g.varint();
                int f2 = (int)(k2 >> 3), w2 = (int)(k2 & 7);
                if (f2 == 1 && w2 == 2) m.nodes.push_back(parse_node(g.sub()));
                else if (f2 == 5 && w2 == 2) { Tensor t = parse_tensor(g.sub()); m.init[t.name] = std::move(t); }
                else g.skip(w2);
            }
        } else r.skip(wt);
    }
    WB_REQUIRE(have_graph, WB_EIO, "%s is not an ONNX ModelProto (no graph)", path.c_str());
    // torch.onnx de-duplicates identical initializers (all-zero biases, all-one LayerNorm weights of an untrained
    // model; it can happen to trained tensors too): one copy stays an initializer, every other name becomes the output
    // of an Identity node fed by it.  Resolve those aliases (chains included) so that lookups by name see them.
    for (bool grew = true; grew;) {
        grew = false;
        for (const Node& nd : m.nodes) {
            if (nd.op != "Identity" || nd.in.size() != 1 || nd.out.size() != 1 || m.init.count(nd.out[0])) continue;
            auto it = m.init.find(nd.in[0]);
            if (it == m.init.end()) continue;
            Tensor alias = it->second;
            alias.name = nd.out[0];
            m.init[nd.out[0]] = std::move(alias);
            grew = true;
        }
    }
}

static float half_to_float(uint16_t h) {
    uint32_t s = (uint32_t)(h & 0x8000) << 16, e = (h >> 10) & 0x1F, f = h & 0x3FF, out;
    if (e == 0) {
        if (f == 0) out = s;
        else { int sh = 0; while (!(f & 0x400)) { f <<= 1; ++sh; } f &= 0x3FF; out = s | ((uint32_t)(113 - sh) << 23) | (f << 13); }
    } else if (e == 31) out = s | 0x7F800000u | (f << 13);
    else out = s | ((e + 112) << 23) | (f << 13);
    float v;
    std::memcpy(&v, &out, 4);
    return v;
}

std::vector<float> to_f32(const Tensor& t) {
    WB_REQUIRE(t.data_location == 0, WB_EIO, "ONNX initializer %s uses external data (not supported)", t.name.c_str());
    const size_t n = (size_t)t.numel();
    std::vector<float> v(n);
    if (t.dtype == 1) {
        if (t.raw_len) { WB_REQUIRE(t.raw_len == n * 4, WB_EIO, "ONNX initializer %s: size mismatch", t.name.c_str()); std::memcpy(v.data(), t.raw, n * 4); }
        else { WB_REQUIRE(t.float_data.size() == n, WB_EIO, "ONNX initializer %s: size mismatch", t.name.c_str()); v = t.float_data; }
    } else if (t.dtype == 10 || t.dtype == 16) {
        WB_REQUIRE(t.raw_len == n * 2, WB_EIO, "ONNX initializer %s: size mismatch", t.name.c_str());
        for (size_t i = 0; i < n; ++i) {
            uint16_t h;
            std::memcpy(&h, t.raw + 2 * i, 2);
            if (t.dtype == 10) v[i] = half_to_float(h);
            else { uint32_t u = (uint32_t)h << 16; std::memcpy(&v[i], &u, 4); }
        }
    } else WB_THROW(WB_EIO, "ONNX initializer %s has unsupported data_type %d", t.name.c_str(), t.dtype);
    return v;
}

// find an initializer by its HF parameter path, whatever module prefix the exporter kept
const Tensor* find_named(const Model& m, const std::string& key, const std::vector<std::string>& prefixes) {
    for (const auto& p : prefixes) {
        auto it = m.init.find(p + key);
        if (it != m.init.end()) return &it->second;
    }
    return nullptr;
}

// anonymous nn.Linear weights: 2-D initializers consumed by MatMul nodes, in node order.  Each is resolved to its HF
// module through the Add that consumes the MatMul output and that Add's NAMED bias initializer ("...q_proj.bias");
// `key` stays empty for the bias-less ones (k_proj, the tied proj_out).
struct AnonLinear {
    const Tensor* w = nullptr;
    std::string out, key;       // MatMul output tensor; HF module path relative to the encoder/decoder ("layers.0.self_attn.q_proj")
    bool used = false;
};
std::vector<AnonLinear> matmul_weights(const Model& m, const std::vector<std::string>& prefixes) {
    std::vector<AnonLinear> out;
    for (const auto& n : m.nodes) {
        if (n.op != "MatMul" || n.in.size() != 2 || n.out.empty()) continue;
        for (const auto& i : n.in) {
            auto it = m.init.find(i);
            if (it != m.init.end() && it->second.dims.size() == 2) {
                AnonLinear a;
                a.w = &it->second;
                a.out = n.out[0];
                out.push_back(a);
            }
        }
    }
    for (const auto& n : m.nodes) {
        if (n.op != "Add" || n.in.size() != 2) continue;
        for (int side = 0; side < 2; ++side) {
            const std::string& b = n.in[side];
            const std::string& x = n.in[1 - side];
            if (b.size() < 6 || b.compare(b.size() - 5, 5, ".bias") != 0 || m.init.find(b) == m.init.end()) continue;
            std::string key = b.substr(0, b.size() - 5);
            for (const auto& p : prefixes)
                if (!p.empty() && key.compare(0, p.size(), p) == 0) { key = key.substr(p.size()); break; }
            for (auto& a : out)
                if (a.out == x) {
                    WB_REQUIRE(a.key.empty() || a.key == key, WB_EINVAL, "ONNX: MatMul weight %s feeds two different biases (%s, %s)",
                               a.w->name.c_str(), a.key.c_str(), key.c_str());
                    a.key = key;
                }
        }
    }
    return out;
}

}  // namespace wbonnx

// Fills `host` with HF-named f32 tensors (same keys weights.cpp expects) from an optimum export dir.
void onnx_load_dir(const std::string& dir, const wb_model_cfg& c, std::map<std::string, std::vector<float>>& host) {
    using namespace wbonnx;
    Model enc, dec;
    load(dir + "/encoder_model.onnx", enc);
    load(dir + "/decoder_model.onnx", dec);
    const int64_t d = c.d_model, f = c.ffn_dim;
    auto take_named = [&](const Model& m, const std::vector<std::string>& pre, const std::string& key, const std::string& hf,
                          std::vector<int64_t> shape) {
        const Tensor* t = find_named(m, key, pre);
        WB_REQUIRE(t != nullptr, WB_EINVAL, "ONNX export misses initializer %s", key.c_str());
        WB_REQUIRE(t->dims == shape, WB_EINVAL, "ONNX initializer %s has an unexpected shape", t->name.c_str());
        host[hf] = to_f32(*t);
    };
    struct Lin { std::string key; int64_t out, in; };
    auto take_linears = [&](const Model& m, const std::vector<std::string>& pre, const std::string& hf_prefix, const std::vector<Lin>& lins,
                            bool tied_tail) {
        // (a) exports that kept nn.Linear names [out,in]; (b) anonymous transposed MatMul weights, identified by the named
        // bias their output is added to; the bias-less ones (k_proj) must be the only unidentified MatMul between their
        // identified neighbours in node order -- anything else is ambiguous and fails instead of loading silently wrong
        std::vector<AnonLinear> anon = matmul_weights(m, pre);
        auto by_key = [&](const std::string& key) -> int {
            int found = -1;
            for (size_t i = 0; i < anon.size(); ++i)
                if (anon[i].key == key) {
                    WB_REQUIRE(found < 0, WB_EINVAL, "ONNX: two MatMul weights resolve to %s", key.c_str());
                    found = (int)i;
                }
            return found;
        };
        auto take_anon = [&](int idx, const Lin& L) {
            AnonLinear& a = anon[(size_t)idx];
            WB_REQUIRE(!a.used, WB_EINVAL, "ONNX MatMul weight %s would be used twice (at %s)", a.w->name.c_str(), L.key.c_str());
            const Tensor* t = a.w;
            WB_REQUIRE(t->dims[0] == L.in && t->dims[1] == L.out, WB_EINVAL,
                       "ONNX MatMul weight %s is [%lld,%lld] but %s needs [%lld,%lld] (in,out): graph order does not match the HF module order",
                       t->name.c_str(), (long long)t->dims[0], (long long)t->dims[1], L.key.c_str(), (long long)L.in, (long long)L.out);
            std::vector<float> w = to_f32(*t), wt((size_t)(L.out * L.in));
            for (int64_t i = 0; i < L.in; ++i)
                for (int64_t o = 0; o < L.out; ++o) wt[(size_t)(o * L.in + i)] = w[(size_t)(i * L.out + o)];
            host[hf_prefix + L.key + ".weight"] = std::move(wt);
            a.used = true;
        };
        std::vector<int> slot(lins.size(), -2);                      // -2: named weight, -1: to be placed by position, >= 0: anon index
        for (size_t li = 0; li < lins.size(); ++li) {
            const Lin& L = lins[li];
            if (const Tensor* t = find_named(m, L.key + ".weight", pre)) {
                WB_REQUIRE(t->dims.size() == 2 && t->dims[0] == L.out && t->dims[1] == L.in, WB_EINVAL, "ONNX initializer %s has an unexpected shape", t->name.c_str());
                host[hf_prefix + L.key + ".weight"] = to_f32(*t);
                continue;
            }
            slot[li] = by_key(L.key);
        }
        for (size_t li = 0; li < lins.size(); ++li)
            if (slot[li] >= 0) take_anon(slot[li], lins[li]);
        for (size_t li = 0; li < lins.size(); ++li) {
            if (slot[li] != -1) continue;
            // bias-less linear: the one unidentified MatMul after the previous linear's and before the next linear's weight
            int lo = -1, hi = (int)anon.size();
            for (size_t j = li; j-- > 0;) if (slot[j] >= 0) { lo = slot[j]; break; }
            for (size_t j = li + 1; j < lins.size(); ++j) if (slot[j] >= 0) { hi = slot[j]; break; }
            int pick = -1, n_free = 0;
            for (int i = lo + 1; i < hi; ++i)
                if (!anon[(size_t)i].used && anon[(size_t)i].key.empty() && anon[(size_t)i].w->dims[0] == lins[li].in && anon[(size_t)i].w->dims[1] == lins[li].out) {
                    if (pick < 0) pick = i;
                    ++n_free;
                }
            // several bias-less linears in a row (an export without any bias) fall back to graph order, one at a time
            size_t run = 1;
            for (size_t j = li + 1; j < lins.size() && slot[j] == -1; ++j) ++run;
            WB_REQUIRE(pick >= 0 && (size_t)n_free == run, WB_EINVAL,
                       "ONNX: cannot place the weight of %s: %d candidate MatMul weights between its neighbours, %zu expected "
                       "(graph order does not match the HF module order)", lins[li].key.c_str(), n_free, run);
            take_anon(pick, lins[li]);
            slot[li] = pick;
        }
        if (tied_tail) {
            for (auto& a : anon)
                if (!a.used && a.key.empty() && a.w->dims[0] == d && a.w->dims[1] == c.vocab) { a.used = true; break; }
        }
        size_t left = 0;
        for (const auto& a : anon) left += a.used ? 0 : 1;
        WB_REQUIRE(left == 0, WB_EINVAL, "ONNX graph has %zu unassigned MatMul weights", left);
    };

    {   // ---------------- encoder_model.onnx ----------------
        const std::vector<std::string> pre = {"model.encoder.", "encoder.", ""};
        const std::string hp = "model.encoder.";
        take_named(enc, pre, "conv1.weight", hp + "conv1.weight", {d, c.n_mels, 3});
        take_named(enc, pre, "conv1.bias", hp + "conv1.bias", {d});
        take_named(enc, pre, "conv2.weight", hp + "conv2.weight", {d, d, 3});
        take_named(enc, pre, "conv2.bias", hp + "conv2.bias", {d});
        if (find_named(enc, "embed_positions.weight", pre)) {
            take_named(enc, pre, "embed_positions.weight", hp + "embed_positions.weight", {c.n_audio_ctx, d});
        } else {
            // the encoder adds the whole position table to the conv output, so torch.onnx folds `embed_positions.weight`
            // into an anonymous `onnx::Add_<n>` initializer: take THE [n_audio_ctx, d] initializer that feeds an Add
            const Tensor* pos = nullptr;
            for (const auto& n : enc.nodes) {
                if (n.op != "Add") continue;
                for (const auto& i : n.in) {
                    auto it = enc.init.find(i);
                    if (it == enc.init.end() || it->second.dims != std::vector<int64_t>{c.n_audio_ctx, d}) continue;
                    WB_REQUIRE(pos == nullptr || pos == &it->second, WB_EINVAL, "ONNX: two candidates for the encoder position table (%s, %s)",
                               pos->name.c_str(), it->second.name.c_str());
                    pos = &it->second;
                }
            }
            WB_REQUIRE(pos != nullptr, WB_EINVAL, "ONNX export misses initializer embed_positions.weight");
            host[hp + "embed_positions.weight"] = to_f32(*pos);
        }
        std::vector<Lin> lins;
        for (int i = 0; i < c.enc_layers; ++i) {
            const std::string L = "layers." + std::to_string(i) + ".";
            for (const char* n : {"self_attn_layer_norm", "final_layer_norm"}) {
                take_named(enc, pre, L + n + ".weight", hp + L + n + ".weight", {d});
                take_named(enc, pre, L + n + ".bias", hp + L + n + ".bias", {d});
            }
            for (const char* n : {"self_attn.q_proj", "self_attn.v_proj", "self_attn.out_proj"}) take_named(enc, pre, L + n + ".bias", hp + L + n + ".bias", {d});
            take_named(enc, pre, L + "fc1.bias", hp + L + "fc1.bias", {f});
            take_named(enc, pre, L + "fc2.bias", hp + L + "fc2.bias", {d});
            lins.push_back({L + "self_attn.q_proj", d, d}); lins.push_back({L + "self_attn.k_proj", d, d});
            lins.push_back({L + "self_attn.v_proj", d, d}); lins.push_back({L + "self_attn.out_proj", d, d});
            lins.push_back({L + "fc1", f, d}); lins.push_back({L + "fc2", d, f});
        }
        take_named(enc, pre, "layer_norm.weight", hp + "layer_norm.weight", {d});
        take_named(enc, pre, "layer_norm.bias", hp + "layer_norm.bias", {d});
        take_linears(enc, pre, hp, lins, false);
    }
    {   // ---------------- decoder_model.onnx ----------------
        const std::vector<std::string> pre = {"model.decoder.", "decoder.", ""};
        const std::string hp = "model.decoder.";
        take_named(dec, pre, "embed_tokens.weight", hp + "embed_tokens.weight", {c.vocab, d});
        take_named(dec, pre, "embed_positions.weight", hp + "embed_positions.weight", {c.n_text_ctx, d});
        std::vector<Lin> lins;
        for (int i = 0; i < c.dec_layers; ++i) {
            const std::string L = "layers." + std::to_string(i) + ".";
            for (const char* n : {"self_attn_layer_norm", "encoder_attn_layer_norm", "final_layer_norm"}) {
                take_named(dec, pre, L + n + ".weight", hp + L + n + ".weight", {d});
                take_named(dec, pre, L + n + ".bias", hp + L + n + ".bias", {d});
            }
            for (const char* a : {"self_attn", "encoder_attn"})
                for (const char* n : {"q_proj", "v_proj", "out_proj"})
                    take_named(dec, pre, L + a + "." + n + ".bias", hp + L + a + "." + n + ".bias", {d});
            take_named(dec, pre, L + "fc1.bias", hp + L + "fc1.bias", {f});
            take_named(dec, pre, L + "fc2.bias", hp + L + "fc2.bias", {d});
            for (const char* a : {"self_attn", "encoder_attn"})
                for (const char* n : {"q_proj", "k_proj", "v_proj", "out_proj"}) lins.push_back({L + a + "." + n, d, d});
            lins.push_back({L + "fc1", f, d}); lins.push_back({L + "fc2", d, f});
        }
        take_named(dec, pre, "layer_norm.weight", hp + "layer_norm.weight", {d});
        take_named(dec, pre, "layer_norm.bias", hp + "layer_norm.bias", {d});
        take_linears(dec, pre, hp, lins, true);
    }
}

// Host-only test hook: load <dir> through the ONNX path and return one tensor by HF name.
extern "C" int wb_onnx_read_tensor(const char* dir, const wb_model_cfg* cfg, const char* name, float* out, int64_t n) {
    try {
        WB_REQUIRE(dir && cfg && name && out, WB_EINVAL, "null argument");
        std::map<std::string, std::vector<float>> host;
        onnx_load_dir(dir, *cfg, host);
        auto it = host.find(name);
        WB_REQUIRE(it != host.end(), WB_EINVAL, "ONNX export has no tensor %s", name);
        WB_REQUIRE((int64_t)it->second.size() == n, WB_EINVAL, "tensor %s has %zu elements, caller asked for %lld", name, it->second.size(), (long long)n);
        std::memcpy(out, it->second.data(), sizeof(float) * (size_t)n);
        return WB_OK;
    } catch (const WbError& e) {
        wb_set_error(e.what());
        return e.code;
    } catch (const std::exception& e) {
        wb_set_error(e.what());
        return WB_EINVAL;
    }
}
