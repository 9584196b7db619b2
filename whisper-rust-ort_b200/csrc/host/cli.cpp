// cli.cpp — the drop-in command line (replaces Args + main of /root/reference/src/main.rs:23-86,
// 1065-1271 and transcribe_longform_chunked :834-1008).  Same flags, same three output files, same
// stdout lines; ORT-only knobs are accepted and echoed in `config_used` but have no effect.  The
// per-file loop is kept serial like the reference by default (so per-file latency means the same
// thing); the parallelism the reference gets from rayon over chunks is the GPU batch dimension here.
// Batch scheduler (BASELINE.json north_star (4)): --gpus N deals the file groups to N worker processes,
// one per GPU (clips are independent: no collective, the parent gathers the per-file rows on the host);
// --in-flight S keeps S groups in flight per GPU, each in its own wb_ctx (stream, caches, graphs).
#include <sys/stat.h>
#include <sys/wait.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <future>
#include <chrono>
#include <cmath>
#include <cstring>
#include <dirent.h>
#include <fstream>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "../common.h"
#include "json.h"

namespace wbtext {
std::string stitch_texts(const std::vector<std::string>& chunks);
std::string csv_field(const std::string& s);
}

namespace {

using Clock = std::chrono::steady_clock;
double since(Clock::time_point t0) { return std::chrono::duration<double>(Clock::now() - t0).count(); }

// WB_CLI_TRACE=1: milestones with seconds since process start on stderr (where does a run spend its wall clock?)
const Clock::time_point g_t0 = Clock::now();
bool trace_on() { static const bool on = std::getenv("WB_CLI_TRACE") != nullptr; return on; }
#define TRACE(...)                                                        \
    do {                                                                  \
        if (trace_on()) {                                                 \
            fprintf(stderr, "[trace %8.3f s] ", since(g_t0));             \
            fprintf(stderr, __VA_ARGS__);                                 \
            fputc('\n', stderr);                                          \
        }                                                                 \
    } while (0)

struct Args {
    std::string audio_dir = "audio", model_id = "openai/whisper-base", onnx_dir = "whisper-base-with-past";
    std::string language = "en", task = "transcribe";
    size_t max_new_tokens = 128, warmup = 0, limit_files = 0;
    std::string discovery_best_json;
    std::string out_csv = "results/benchmarks/inference_per_file.csv";
    std::string out_json = "results/benchmarks/inference_per_file.json";
    std::string out_summary_json = "results/benchmarks/inference_summary.json";
    size_t intra_op = 0, inter_op = 0;
    bool write_txt = false;
    std::string tokenizer_json;
    bool timestamps = false;
    size_t chunk_parallelism = 0;
    float chunk_length_s = 30.0f, overlap_s = 5.0f;
    // extensions (not in the reference)
    int device = 0;
    std::string precision = "bf16", weights, arch = "base";
    bool arch_given = false;
    uint64_t seed = 0;
    int batch = 32;
    size_t file_batch = 1;      // files transcribed together (1 = the reference's serial per-file loop)
    size_t gpus = 1;            // worker processes, one per GPU: devices device .. device+gpus-1
    size_t in_flight = 1;       // file groups in flight per GPU (one wb_ctx each)
};

struct OrtCfg {      // main.rs:91-100 — echoed only
    size_t intra_op, inter_op;
    std::string execution_mode, graph_opt;
    bool cpu_mem_arena, mem_pattern, allow_spinning;
};

struct UsageError : std::runtime_error { using std::runtime_error::runtime_error; };

const char* USAGE =
    "Usage: whisper_b200_cli [OPTIONS]\n\n"
    "Options (identical to whisper_ort_bench, /root/reference/src/main.rs:23-86):\n"
    "      --audio-dir <DIR>              [default: audio]\n"
    "      --model-id <ID>                [default: openai/whisper-base]\n"
    "      --onnx-dir <DIR>               [default: whisper-base-with-past]\n"
    "      --language <LANG>              [default: en]\n"
    "      --task <TASK>                  [default: transcribe]\n"
    "      --max-new-tokens <N>           [default: 128]\n"
    "      --warmup <N>                   [default: 0]\n"
    "      --limit-files <N>              [default: 0]\n"
    "      --discovery-best-json <PATH>   [default: ]\n"
    "      --out-csv <PATH>               [default: results/benchmarks/inference_per_file.csv]\n"
    "      --out-json <PATH>              [default: results/benchmarks/inference_per_file.json]\n"
    "      --out-summary-json <PATH>      [default: results/benchmarks/inference_summary.json]\n"
    "      --intra-op <N>                 accepted, echoed, no effect on the GPU path\n"
    "      --inter-op <N>                 accepted, echoed, no effect on the GPU path\n"
    "      --write-txt\n"
    "      --tokenizer-json <PATH>\n"
    "      --timestamps\n"
    "      --chunk-parallelism <N>        accepted; chunks are batched on the GPU instead\n"
    "      --chunk-length-s <S>           [default: 30]\n"
    "      --overlap-s <S>                [default: 5]\n"
    "B200 extensions:\n"
    "      --device <N>  --precision <bf16|fp32>  --batch <N>  --weights <file.wb200>  --arch <base|large-v3|toy>  --seed <N>\n"
    "      --file-batch <N>               files whose chunks share GPU batches [default: 1 = serial like the reference];\n"
    "                                     every file of a group is charged the group's preprocess/model/decode time\n"
    "      --gpus <N>                     shard the file groups over N GPUs, one worker process per GPU [default: 1]\n"
    "      --in-flight <S>                file groups in flight per GPU, one context each [default: 1]\n"
    "  -h, --help\n";

Args parse_args(int argc, const char* const* argv) {
    Args a;
    auto need = [&](int& i, const std::string& flag, const std::string& inline_v, bool has_inline) -> std::string {
        if (has_inline) return inline_v;
        if (i + 1 >= argc) throw UsageError("error: a value is required for '" + flag + "' but none was supplied");
        return argv[++i];
    };
    auto to_usize = [&](const std::string& v, const std::string& flag) -> size_t {
        char* end = nullptr;
        if (v.empty() || v[0] == '-') throw UsageError("error: invalid value '" + v + "' for '" + flag + "'");
        unsigned long long r = std::strtoull(v.c_str(), &end, 10);
        if (*end) throw UsageError("error: invalid value '" + v + "' for '" + flag + "'");
        return (size_t)r;
    };
    auto to_f32 = [&](const std::string& v, const std::string& flag) -> float {
        char* end = nullptr;
        float r = std::strtof(v.c_str(), &end);
        if (v.empty() || *end) throw UsageError("error: invalid value '" + v + "' for '" + flag + "'");
        return r;
    };
    for (int i = 1; i < argc; ++i) {
        std::string tok = argv[i], flag = tok, iv;
        bool has_inline = false;
        size_t eq = tok.find('=');
        if (tok.rfind("--", 0) == 0 && eq != std::string::npos) { flag = tok.substr(0, eq); iv = tok.substr(eq + 1); has_inline = true; }
        if (flag == "-h" || flag == "--help") throw UsageError("");
        else if (flag == "--audio-dir") a.audio_dir = need(i, flag, iv, has_inline);
        else if (flag == "--model-id") a.model_id = need(i, flag, iv, has_inline);
        else if (flag == "--onnx-dir") a.onnx_dir = need(i, flag, iv, has_inline);
        else if (flag == "--language") a.language = need(i, flag, iv, has_inline);
        else if (flag == "--task") a.task = need(i, flag, iv, has_inline);
        else if (flag == "--max-new-tokens") a.max_new_tokens = to_usize(need(i, flag, iv, has_inline), flag);
        else if (flag == "--warmup") a.warmup = to_usize(need(i, flag, iv, has_inline), flag);
        else if (flag == "--limit-files") a.limit_files = to_usize(need(i, flag, iv, has_inline), flag);
        else if (flag == "--discovery-best-json") a.discovery_best_json = need(i, flag, iv, has_inline);
        else if (flag == "--out-csv") a.out_csv = need(i, flag, iv, has_inline);
        else if (flag == "--out-json") a.out_json = need(i, flag, iv, has_inline);
        else if (flag == "--out-summary-json") a.out_summary_json = need(i, flag, iv, has_inline);
        else if (flag == "--intra-op") a.intra_op = to_usize(need(i, flag, iv, has_inline), flag);
        else if (flag == "--inter-op") a.inter_op = to_usize(need(i, flag, iv, has_inline), flag);
        else if (flag == "--write-txt") a.write_txt = true;
        else if (flag == "--tokenizer-json") a.tokenizer_json = need(i, flag, iv, has_inline);
        else if (flag == "--timestamps") a.timestamps = true;
        else if (flag == "--chunk-parallelism") a.chunk_parallelism = to_usize(need(i, flag, iv, has_inline), flag);
        else if (flag == "--chunk-length-s") a.chunk_length_s = to_f32(need(i, flag, iv, has_inline), flag);
        else if (flag == "--overlap-s") a.overlap_s = to_f32(need(i, flag, iv, has_inline), flag);
        else if (flag == "--device") a.device = (int)to_usize(need(i, flag, iv, has_inline), flag);
        else if (flag == "--precision") a.precision = need(i, flag, iv, has_inline);
        else if (flag == "--weights") a.weights = need(i, flag, iv, has_inline);
        else if (flag == "--arch") { a.arch = need(i, flag, iv, has_inline); a.arch_given = true; }
        else if (flag == "--seed") a.seed = to_usize(need(i, flag, iv, has_inline), flag);
        else if (flag == "--batch") a.batch = (int)to_usize(need(i, flag, iv, has_inline), flag);
        else if (flag == "--file-batch") a.file_batch = std::max<size_t>(1, to_usize(need(i, flag, iv, has_inline), flag));
        else if (flag == "--gpus") a.gpus = std::max<size_t>(1, to_usize(need(i, flag, iv, has_inline), flag));
        else if (flag == "--in-flight") a.in_flight = std::max<size_t>(1, to_usize(need(i, flag, iv, has_inline), flag));
        else throw UsageError("error: unexpected argument '" + tok + "' found");
    }
    return a;
}

bool is_file(const std::string& p) { struct stat st; return ::stat(p.c_str(), &st) == 0 && S_ISREG(st.st_mode); }
bool is_dir(const std::string& p) { struct stat st; return ::stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode); }
std::string parent_of(const std::string& p) {
    size_t k = p.find_last_of('/');
    return k == std::string::npos ? std::string() : (k == 0 ? std::string("/") : p.substr(0, k));
}
void create_dir_all(const std::string& p) {
    if (p.empty() || is_dir(p)) return;
    create_dir_all(parent_of(p));
    if (::mkdir(p.c_str(), 0777) != 0 && !is_dir(p)) WB_THROW(WB_EIO, "cannot create directory %s", p.c_str());
}
std::string join(const std::string& a, const std::string& b) { return a.empty() ? b : (a.back() == '/' ? a + b : a + "/" + b); }
std::string read_file(const std::string& p) {
    std::ifstream f(p, std::ios::binary);
    WB_REQUIRE(f.good(), WB_EIO, "cannot read %s", p.c_str());
    std::stringstream ss;
    ss << f.rdbuf();
    return ss.str();
}
void write_file(const std::string& p, const std::string& s) {
    std::ofstream f(p, std::ios::binary);
    WB_REQUIRE(f.good(), WB_EIO, "cannot write %s", p.c_str());
    f << s;
}

OrtCfg suggested_optimum_cfg() {                                                    // main.rs:108-122
    unsigned cpu = std::thread::hardware_concurrency();
    if (cpu == 0) cpu = 8;
    return OrtCfg{std::min<size_t>(cpu, 16), 1, "SEQUENTIAL", "ENABLE_ALL", true, true, true};
}

OrtCfg load_best_cfg_from_discovery(const std::string& path) {                      // main.rs:124-167
    wbjson::Value outer = wbjson::parse(read_file(path));
    const wbjson::Value& best = outer["best"];
    auto lower_trim = [](std::string s) {
        size_t b = s.find_first_not_of(" \t\r\n"), e = s.find_last_not_of(" \t\r\n");
        s = b == std::string::npos ? "" : s.substr(b, e - b + 1);
        for (auto& c : s) c = (char)std::tolower((unsigned char)c);
        return s;
    };
    auto get_bool = [&](const char* k, bool def) {
        const wbjson::Value& v = best[k];
        if (v.type == wbjson::Value::Bool) return v.b;
        if (v.type == wbjson::Value::Num) return v.is_int ? (long long)v.n != 0 : false;
        if (v.type == wbjson::Value::Str) { std::string s = lower_trim(v.s); return s == "1" || s == "true" || s == "yes" || s == "y" || s == "on"; }
        return def;
    };
    auto get_usize = [&](const char* k, size_t def) {
        const wbjson::Value& v = best[k];
        if (v.type == wbjson::Value::Num) return (v.is_int && v.n >= 0) ? (size_t)v.n : def;
        if (v.type == wbjson::Value::Str) { char* e = nullptr; unsigned long long r = std::strtoull(v.s.c_str(), &e, 10); return (v.s.empty() || *e) ? def : (size_t)r; }
        return def;
    };
    auto get_string = [&](const char* k, const char* def) {
        const wbjson::Value& v = best[k];
        return v.type == wbjson::Value::Str ? v.s : std::string(def);
    };
    OrtCfg fb = suggested_optimum_cfg();
    return OrtCfg{get_usize("intra_op", fb.intra_op), get_usize("inter_op", 1), get_string("execution_mode", "SEQUENTIAL"),
                  get_string("graph_opt", "ENABLE_ALL"), get_bool("cpu_mem_arena", true), get_bool("mem_pattern", true),
                  get_bool("allow_spinning", true)};
}

std::string cfg_json(const OrtCfg& c, int indent) {      // struct field order (serde derive), pretty
    std::string in((size_t)indent + 2, ' '), end((size_t)indent, ' ');
    std::string s = "{\n";
    s += in + "\"intra_op\": " + std::to_string(c.intra_op) + ",\n";
    s += in + "\"inter_op\": " + std::to_string(c.inter_op) + ",\n";
    s += in + "\"execution_mode\": " + wbjson::escape(c.execution_mode) + ",\n";
    s += in + "\"graph_opt\": " + wbjson::escape(c.graph_opt) + ",\n";
    s += in + "\"cpu_mem_arena\": " + (c.cpu_mem_arena ? "true" : "false") + ",\n";
    s += in + "\"mem_pattern\": " + (c.mem_pattern ? "true" : "false") + ",\n";
    s += in + "\"allow_spinning\": " + (c.allow_spinning ? "true" : "false") + "\n";
    return s + end + "}";
}

struct Tok {
    wb_tokenizer* t = nullptr;
    std::string path;
    ~Tok() { if (t) wb_tokenizer_free(t); }
};

void load_tok(Tok& tk, const std::string& p) {
    if (wb_tokenizer_load(&tk.t, p.c_str()) != WB_OK) WB_THROW(WB_EIO, "%s", wb_last_error());
    tk.path = p;
}

void resolve_tokenizer(const Args& a, Tok& tk) {                                     // main.rs:574-635
    auto trim = [](const std::string& s) {
        size_t b = s.find_first_not_of(" \t\r\n"), e = s.find_last_not_of(" \t\r\n");
        return b == std::string::npos ? std::string() : s.substr(b, e - b + 1);
    };
    std::string tj = trim(a.tokenizer_json);
    if (!tj.empty()) {
        WB_REQUIRE(is_file(tj), WB_EIO, "tokenizer_json not found: %s", tj.c_str());
        load_tok(tk, tj);
        return;
    }
    for (const std::string& p : {join(a.onnx_dir, "tokenizer.json"), join(a.model_id, "tokenizer.json")})
        if (is_file(p)) { load_tok(tk, p); return; }
    size_t slash = a.model_id.find('/');
    if (slash != std::string::npos) {
        std::string org = a.model_id.substr(0, slash), name = a.model_id.substr(slash + 1);
        if (!org.empty() && !name.empty()) {
            const char* hf = std::getenv("HF_HOME");
            const char* home = std::getenv("HOME");
            std::string base = hf ? std::string(hf) : join(home ? home : ".", ".cache/huggingface");
            std::string snaps = join(join(join(base, "hub"), "models--" + org + "--" + name), "snapshots");
            if (is_dir(snaps)) {
                std::string best;
                time_t best_m = 0;
                bool have = false;
                if (DIR* d = ::opendir(snaps.c_str())) {
                    while (dirent* e = ::readdir(d)) {
                        std::string nm = e->d_name;
                        if (nm == "." || nm == "..") continue;
                        std::string p = join(join(snaps, nm), "tokenizer.json");
                        struct stat st;
                        if (is_file(p) && ::stat(join(snaps, nm).c_str(), &st) == 0 && (!have || st.st_mtime > best_m)) {
                            best = p; best_m = st.st_mtime; have = true;
                        }
                    }
                    ::closedir(d);
                }
                if (have) load_tok(tk, best);
            }
        }
    }
}

struct GenCfg { std::vector<int64_t> suppress, begin_suppress; };
GenCfg load_generation_cfg(const std::string& path) {                               // main.rs:650-657
    GenCfg g;
    if (!is_file(path)) return g;
    wbjson::Value v = wbjson::parse(read_file(path));
    for (const auto& x : v["suppress_tokens"].arr()) g.suppress.push_back((int64_t)x.num());
    for (const auto& x : v["begin_suppress_tokens"].arr()) g.begin_suppress.push_back((int64_t)x.num());
    return g;
}

struct Timing { double preprocess_s = 0, model_only_s = 0, decode_s = 0, end_to_end_s = 0; };

}  // namespace

// The architecture of an export directory from the HF `config.json` optimum writes next to the .onnx files.  The reference
// needs no such thing (the ONNX graphs carry their shapes, so `--onnx-dir` may hold any Whisper size); here the shapes size
// the device buffers.  Returns WB_OK and fills the nine architecture fields, or an error when a field is missing.
extern "C" int wb_cfg_from_hf_config(const char* config_json_path, wb_model_cfg* cfg) {
    try {
        WB_REQUIRE(config_json_path && cfg, WB_EINVAL, "null argument");
        WB_REQUIRE(is_file(config_json_path), WB_EIO, "no such file: %s", config_json_path);
        wbjson::Value v = wbjson::parse(read_file(config_json_path));
        auto field = [&](const char* key) -> int32_t {
            const wbjson::Value& x = v[key];
            WB_REQUIRE(!x.is_null() && x.num() >= 1 && x.num() == (double)(int32_t)x.num(), WB_EINVAL, "%s: no usable \"%s\"", config_json_path, key);
            return (int32_t)x.num();
        };
        wb_model_cfg c = *cfg;
        c.n_mels = field("num_mel_bins"); c.d_model = field("d_model"); c.n_heads = field("encoder_attention_heads");
        c.ffn_dim = field("encoder_ffn_dim"); c.enc_layers = field("encoder_layers"); c.dec_layers = field("decoder_layers");
        c.vocab = field("vocab_size"); c.n_audio_ctx = field("max_source_positions"); c.n_text_ctx = field("max_target_positions");
        WB_REQUIRE(field("decoder_attention_heads") == c.n_heads && field("decoder_ffn_dim") == c.ffn_dim, WB_EINVAL,
                   "%s: encoder and decoder widths differ (not a Whisper layout)", config_json_path);
        WB_REQUIRE(c.d_model % c.n_heads == 0, WB_EINVAL, "%s: d_model %d is not a multiple of %d heads", config_json_path, c.d_model, c.n_heads);
        *cfg = c;
        return WB_OK;
    } catch (const WbError& e) {
        wb_set_error(e.what());
        return e.code;
    } catch (const std::exception& e) {
        wb_set_error(e.what());
        return WB_EINVAL;
    }
}

namespace {

#define CK(call)                                                  \
    do {                                                          \
        int _rc = (call);                                         \
        if (_rc != WB_OK) throw WbError(_rc, wb_last_error());    \
    } while (0)

// transcribe_longform_chunked, main.rs:834-1008, for a GROUP of files: one log-mel launch over all of them
// (the clamp maximum stays per file, quirk Q1), all their chunks packed into GPU batches of max_batch, then
// per-file detokenise + stitch.  With one file per group this is exactly the reference's per-file call.
// `pcm` holds the files back to back (offs[i] .. offs[i+1]), in pinned memory when it comes from the group loader.
std::vector<std::string> transcribe(wb_ctx* ctx, int max_batch, const float* pcm, const std::vector<int64_t>& offs, const Args& a,
                                    const wb_tokenizer* tok, const GenCfg& gen, Timing& t) {
    const size_t n_files = offs.size() - 1;
    auto t0 = Clock::now();
    int64_t sp[5];
    CK(wb_host_special_tokens(tok, a.language.c_str(), a.task.c_str(), sp));
    std::vector<int64_t> prompt = {sp[0], sp[2], sp[3]};
    if (!a.timestamps) prompt.push_back(sp[4]);
    const int64_t eot = sp[1];
    const int64_t chunk_len = (int64_t)std::lround(a.chunk_length_s * 16000.0f);     // main.rs:859-861
    const int64_t overlap = (int64_t)std::lround(a.overlap_s * 16000.0f);
    const int64_t step = std::max<int64_t>(chunk_len > overlap ? chunk_len - overlap : 0, 1);

    auto tp0 = Clock::now();
    int n_chunks = 0;
    CK(wb_log_mel(ctx, pcm, offs.data(), (int)n_files, chunk_len, step, nullptr, nullptr, &n_chunks));
    std::vector<int32_t> chunk_file((size_t)n_chunks);
    CK(wb_get_chunks(ctx, chunk_file.data(), nullptr, n_chunks));
    t.preprocess_s += since(tp0);

    const int mn = (int)std::max<size_t>(a.max_new_tokens, 1);
    const int stride = (int)prompt.size() + mn;
    std::vector<int64_t> toks((size_t)n_chunks * stride);
    std::vector<int32_t> lens((size_t)n_chunks);
    auto tm0 = Clock::now();
    for (int c0 = 0; c0 < n_chunks; c0 += max_batch) {
        const int B = std::min(max_batch, n_chunks - c0);
        CK(wb_encode(ctx, nullptr, c0, B, nullptr));
        CK(wb_greedy_decode(ctx, B, prompt.data(), (int)prompt.size(), (int)a.max_new_tokens, eot,
                            gen.suppress.data(), (int)gen.suppress.size(), gen.begin_suppress.data(),
                            (int)gen.begin_suppress.size(), toks.data() + (size_t)c0 * stride, lens.data() + c0, nullptr, nullptr));
    }
    t.model_only_s += since(tm0);

    auto td0 = Clock::now();
    std::vector<std::vector<std::string>> texts(n_files);
    for (int c = 0; c < n_chunks; ++c) {                                             // main.rs:925-943
        const int64_t* row = toks.data() + (size_t)c * stride;
        int len = lens[c];
        std::vector<int64_t> g;
        if (len > (int)prompt.size()) g.assign(row + prompt.size(), row + len);
        if (!g.empty() && g.back() == eot) g.pop_back();
        int64_t need = wb_host_decode_tokens(tok, g.data(), (int)g.size(), nullptr, 0);
        WB_REQUIRE(need >= 0, WB_EINVAL, "Tokenizer decode failed");
        std::string text((size_t)need + 1, '\0');
        wb_host_decode_tokens(tok, g.data(), (int)g.size(), &text[0], need + 1);
        text.resize((size_t)need);
        if (text.empty()) text = "[EMPTY]";
        if (text != "[EMPTY]") texts[(size_t)chunk_file[c]].push_back(text);
    }
    t.decode_s += since(td0);
    std::vector<std::string> full;
    for (const auto& tx : texts) full.push_back(wbtext::stitch_texts(tx));
    t.end_to_end_s = since(t0);
    return full;
}

// ---- batch scheduler: groups of files -> worker contexts -> per-file results ----
struct FileResult { bool done = false; double dur = 0, load_s = 0; Timing t; std::string text; };
struct Job {            // shared, read-only
    const Args& args;
    const std::vector<std::string>& files;
    const wb_tokenizer* tok;
    const GenCfg& gen;
    wb_model_cfg mc;
    std::string wpath;
    size_t decode_threads = 1;      // host threads a group loader may use to decode its files (from --intra-op, capped)
};
struct Pcm {
    float* p = nullptr; int64_t n = 0; double dur = 0;
    Pcm() = default; Pcm(const Pcm&) = delete; Pcm& operator=(const Pcm&) = delete;
    ~Pcm() { wb_host_free(p); }
};
// page-locked staging of one group's PCM (grown on demand, reused from group to group)
struct PinnedPcm {
    float* p = nullptr; size_t cap = 0;
    PinnedPcm() = default; PinnedPcm(const PinnedPcm&) = delete; PinnedPcm& operator=(const PinnedPcm&) = delete;
    ~PinnedPcm() { wb_host_free_pinned(p); }
    void reserve(int device, size_t n) {
        if (n <= cap) return;
        wb_host_free_pinned(p); p = nullptr; cap = 0;
        void* q = nullptr;
        CK(wb_host_alloc_pinned(device, sizeof(float) * n, &q));
        p = static_cast<float*>(q); cap = n;
    }
};
// One group of files, decoded and packed back to back: what a worker hands to transcribe().
struct LoadedGroup {
    size_t g = SIZE_MAX, g0 = 0, g1 = 0;          // g == SIZE_MAX: no group left
    std::vector<int64_t> offs;
    std::vector<double> dur, load_s;
    double load_total = 0;
};

// The k-th group of this rank: read + decode + downmix/resample every file (main.rs:207-316), pack into `slot`.
LoadedGroup load_group(const Job& J, int device, const std::vector<std::pair<size_t, size_t>>& groups, size_t rank, size_t world,
                       std::atomic<size_t>& cursor, PinnedPcm& slot) {
    LoadedGroup L;
    const size_t k = cursor.fetch_add(1);
    const size_t g = rank + k * world;
    if (g >= groups.size()) return L;
    L.g = g; L.g0 = groups[g].first; L.g1 = groups[g].second;
    const size_t n = L.g1 - L.g0;
    std::vector<Pcm> au(n);
    L.offs.assign(n + 1, 0); L.dur.resize(n); L.load_s.resize(n);
    // files of a group are independent: decode them on up to decode_threads host threads (a group of one file, the
    // reference's serial loop, stays on this thread); the first failure in file order is the one reported
    std::vector<std::exception_ptr> err(n);
    std::atomic<size_t> next_file{0};
    auto decode_files = [&]() {
        for (size_t i; (i = next_file.fetch_add(1)) < n;) {
            try {
                auto tl0 = Clock::now();
                Pcm& p = au[i];
                CK(wb_host_load_audio_16k_mono(join(J.args.audio_dir, J.files[L.g0 + i]).c_str(), &p.p, &p.n, &p.dur));
                WB_REQUIRE(p.n > 0, WB_EINVAL, "Empty audio");
                L.dur[i] = p.dur;
                L.load_s[i] = since(tl0);
            } catch (...) {
                err[i] = std::current_exception();
            }
        }
    };
    {
        std::vector<std::thread> helpers;
        for (size_t t = 1; t < std::min(J.decode_threads, n); ++t) helpers.emplace_back(decode_files);
        decode_files();
        for (std::thread& t : helpers) t.join();
    }
    for (size_t i = 0; i < n; ++i) {
        if (err[i]) std::rethrow_exception(err[i]);
        L.offs[i + 1] = L.offs[i] + au[i].n;
    }
    auto tc0 = Clock::now();
    slot.reserve(device, (size_t)L.offs.back());
    for (size_t i = 0; i < n; ++i) std::memcpy(slot.p + L.offs[i], au[i].p, sizeof(float) * (size_t)au[i].n);
    const double pack_s = since(tc0) / (double)n;               // the staging copy is part of loading a file
    for (size_t i = 0; i < n; ++i) { L.load_s[i] += pack_s; L.load_total += L.load_s[i]; }
    return L;
}

// One worker = one wb_ctx on `device`: warm up like main.rs:1131-1152, then take groups off the cursor.  A loader
// thread decodes the worker's NEXT group into the other pinned slot while the GPU works on the current one, so with
// one context per GPU the device only waits for the first group of the run.
void run_worker(const Job& J, int device, const std::vector<std::pair<size_t, size_t>>& groups, size_t rank, size_t world,
                std::atomic<size_t>& cursor, std::vector<FileResult>& out) {
    const Args& args = J.args;
    wb_ctx* ctx = nullptr;
    TRACE("worker on gpu %d: creating context", device);
    CK(wb_create(&ctx, device, &J.mc, J.wpath.empty() ? nullptr : J.wpath.c_str()));
    struct Guard { wb_ctx* c; ~Guard() { wb_destroy(c); } } guard{ctx};
    CK(wb_set_load_hint(ctx, (int)args.in_flight));     // the reference's serial loop (1 in flight) gets the latency-oriented kernels
    TRACE("worker on gpu %d: context ready", device);
    PinnedPcm slots[2];
    auto prefetch = [&](int which) {
        return std::async(std::launch::async, [&, which] { return load_group(J, device, groups, rank, world, cursor, slots[which]); });
    };
    std::future<LoadedGroup> next = prefetch(0);                 // overlaps the warm-up
    struct Drain { std::future<LoadedGroup>& f; ~Drain() { if (f.valid()) f.wait(); } } drain{next};   // never leave a loader running
    if (args.warmup > 0) {
        Pcm a0;
        CK(wb_host_load_audio_16k_mono(join(args.audio_dir, J.files[0]).c_str(), &a0.p, &a0.n, &a0.dur));
        WB_REQUIRE(a0.n > 0, WB_EINVAL, "Empty audio");
        // main.rs:1131-1152 warms up on the first file.  With file groups, the warm-up runs that file as many times as a
        // group has files, so the decode graphs of the real batch shape are built here and not inside the first group.
        const int64_t w_chunk = (int64_t)std::lround(args.chunk_length_s * 16000.0f), w_ov = (int64_t)std::lround(args.overlap_s * 16000.0f);
        const int n_ch0 = std::max(1, wb_host_chunk_starts(a0.n, w_chunk, std::max<int64_t>(w_chunk > w_ov ? w_chunk - w_ov : 0, 1), nullptr, 0));
        const size_t copies = std::max<size_t>(1, std::min<size_t>({args.file_batch, (size_t)J.mc.max_batch, (size_t)(J.mc.max_chunks / n_ch0)}));
        std::vector<int64_t> woffs(copies + 1, 0);
        for (size_t i = 0; i < copies; ++i) woffs[i + 1] = woffs[i] + a0.n;
        PinnedPcm wbuf;
        const float* wp = a0.p;
        if (copies > 1) {
            wbuf.reserve(device, (size_t)woffs.back());
            for (size_t i = 0; i < copies; ++i) std::memcpy(wbuf.p + woffs[i], a0.p, sizeof(float) * (size_t)a0.n);
            wp = wbuf.p;
        }
        for (size_t i = 0; i < args.warmup; ++i) { Timing t; transcribe(ctx, J.mc.max_batch, wp, woffs, args, J.tok, J.gen, t); }
        TRACE("worker on gpu %d: warm-up done", device);
    }
    for (int cur = 0;; cur ^= 1) {
        LoadedGroup L = next.get();
        if (L.g == SIZE_MAX) break;
        next = prefetch(cur ^ 1);
        Timing t;
        std::vector<std::string> texts = transcribe(ctx, J.mc.max_batch, slots[cur].p, L.offs, args, J.tok, J.gen, t);
        TRACE("gpu %d: group %zu (%zu files) load %.3f s, preprocess %.3f s, model %.3f s, detokenise %.3f s", device, L.g, L.g1 - L.g0,
              L.load_total, t.preprocess_s, t.model_only_s, t.decode_s);
        for (size_t i = L.g0; i < L.g1; ++i) {
            FileResult& fr = out[i];
            fr.dur = L.dur[i - L.g0]; fr.load_s = L.load_s[i - L.g0]; fr.t = t; fr.text = texts[i - L.g0];
            fr.done = true;
        }
    }
}

// All groups of one rank (= one GPU), `in_flight` workers at a time.  Files stay in listing order inside a
// worker, so --in-flight 1 is the reference's serial loop.
void run_rank(const Job& J, int device, const std::vector<std::pair<size_t, size_t>>& groups, size_t rank, size_t world,
              std::vector<FileResult>& out) {
    std::atomic<size_t> cursor{0};
    const size_t mine = groups.size() > rank ? (groups.size() - rank + world - 1) / world : 0;
    const size_t n_workers = std::max<size_t>(1, std::min(J.args.in_flight, mine));
    if (n_workers == 1) { run_worker(J, device, groups, rank, world, cursor, out); return; }
    std::mutex mu;
    std::string first_error;
    std::vector<std::thread> th;
    for (size_t w = 0; w < n_workers; ++w)
        th.emplace_back([&] {
            try { run_worker(J, device, groups, rank, world, cursor, out); }
            catch (const std::exception& e) {
                std::lock_guard<std::mutex> lk(mu);
                if (first_error.empty()) first_error = e.what();
                cursor.store(groups.size());                  // stop handing out work
            }
        });
    for (auto& t : th) t.join();
    if (!first_error.empty()) WB_THROW(WB_ESTATE, "%s", first_error.c_str());
}

// rows of one worker process, one JSON object per line (f64 printed shortest-round-trip, so nothing is lost)
void write_rows(const std::string& path, const std::vector<FileResult>& rs) {
    std::string s;
    for (size_t i = 0; i < rs.size(); ++i) {
        if (!rs[i].done) continue;
        const FileResult& r = rs[i];
        s += "{\"i\": " + std::to_string(i) + ", \"dur\": " + wbjson::fmt_f64(r.dur) + ", \"load\": " + wbjson::fmt_f64(r.load_s) +
             ", \"pre\": " + wbjson::fmt_f64(r.t.preprocess_s) + ", \"model\": " + wbjson::fmt_f64(r.t.model_only_s) +
             ", \"dec\": " + wbjson::fmt_f64(r.t.decode_s) + ", \"e2e\": " + wbjson::fmt_f64(r.t.end_to_end_s) +
             ", \"text\": " + wbjson::escape(r.text) + "}\n";
    }
    write_file(path, s);
}
void read_rows(const std::string& path, std::vector<FileResult>& rs) {
    std::istringstream in(read_file(path));
    std::string line;
    while (std::getline(in, line)) {
        if (line.empty()) continue;
        wbjson::Value v = wbjson::parse(line);
        const size_t i = (size_t)v["i"].num();
        WB_REQUIRE(i < rs.size(), WB_ESTATE, "worker row %zu out of range", i);
        FileResult& r = rs[i];
        r.dur = v["dur"].num(); r.load_s = v["load"].num();
        r.t.preprocess_s = v["pre"].num(); r.t.model_only_s = v["model"].num(); r.t.decode_s = v["dec"].num(); r.t.end_to_end_s = v["e2e"].num();
        r.text = v["text"].str();
        r.done = true;
    }
}

std::string stat_json(const std::vector<double>& xs, int indent) {                  // keys sorted (serde_json BTreeMap)
    double o[6];
    wb_host_stat_block(xs.data(), (int)xs.size(), o);
    std::string in((size_t)indent + 2, ' '), end((size_t)indent, ' ');
    return "{\n" + in + "\"max\": " + wbjson::fmt_f64(o[4]) + ",\n" + in + "\"mean\": " + wbjson::fmt_f64(o[5]) + ",\n" + in +
           "\"median\": " + wbjson::fmt_f64(o[1]) + ",\n" + in + "\"min\": " + wbjson::fmt_f64(o[0]) + ",\n" + in +
           "\"p90\": " + wbjson::fmt_f64(o[2]) + ",\n" + in + "\"p95\": " + wbjson::fmt_f64(o[3]) + "\n" + end + "}";
}

std::string csv_field(const std::string& s) { return wbtext::csv_field(s); }

std::string fmt_fixed(double v, int prec) { char b[64]; snprintf(b, sizeof(b), "%.*f", prec, v); return b; }
std::string lower(std::string s) { for (auto& c : s) c = (char)std::tolower((unsigned char)c); return s; }

int run(const Args& args) {
    create_dir_all(parent_of(args.out_csv));
    create_dir_all(parent_of(args.out_json));
    create_dir_all(parent_of(args.out_summary_json));

    OrtCfg cfg = !args.discovery_best_json.empty() ? load_best_cfg_from_discovery(args.discovery_best_json) : suggested_optimum_cfg();
    if (args.intra_op > 0) cfg.intra_op = args.intra_op;
    if (args.inter_op > 0) cfg.inter_op = args.inter_op;

    Tok tk;
    resolve_tokenizer(args, tk);
    GenCfg gen = load_generation_cfg(join(args.onnx_dir, "generation_config.json"));

    WB_REQUIRE(is_dir(args.onnx_dir), WB_EIO, "onnx_dir does not exist or is not a directory: %s", args.onnx_dir.c_str());

    // weights (BASELINE.json north_star): --weights / <onnx_dir>/weights.wb200, else the initializers of the
    // reference's own encoder_model.onnx + decoder_model.onnx, else seeded random init of the architecture.
    std::string wpath = args.weights;
    if (wpath.empty() && is_file(join(args.onnx_dir, "weights.wb200"))) wpath = join(args.onnx_dir, "weights.wb200");
    if (wpath.empty() && is_file(join(args.onnx_dir, "encoder_model.onnx")) && is_file(join(args.onnx_dir, "decoder_model.onnx")))
        wpath = args.onnx_dir;
    if (wpath.empty())
        fprintf(stderr, "note: no weights in %s (weights.wb200 or encoder_model.onnx + decoder_model.onnx): "
                        "using seeded random-init %s weights\n", args.onnx_dir.c_str(), args.arch.c_str());
    wb_model_cfg mc;
    CK(wb_default_cfg(&mc, args.arch.c_str()));
    // like the reference, take the model size from the export directory: its config.json wins unless --arch was given
    if (!args.arch_given && is_file(join(args.onnx_dir, "config.json")) &&
        wb_cfg_from_hf_config(join(args.onnx_dir, "config.json").c_str(), &mc) != WB_OK)
        fprintf(stderr, "note: %s is not a Whisper config (%s): keeping --arch %s\n", join(args.onnx_dir, "config.json").c_str(),
                wb_last_error(), args.arch.c_str());
    WB_REQUIRE(args.precision == "bf16" || args.precision == "fp32", WB_EINVAL, "--precision must be bf16 or fp32");
    mc.precision = args.precision == "bf16" ? WB_PREC_BF16 : WB_PREC_FP32;
    mc.max_batch = std::max(1, args.batch);
    mc.max_chunks = std::max(1024, mc.max_batch);
    mc.seed = args.seed;
    // list audio files (main.rs:1111-1128)
    std::vector<std::string> files;
    DIR* d = ::opendir(args.audio_dir.c_str());
    WB_REQUIRE(d != nullptr, WB_EIO, "cannot read audio dir %s", args.audio_dir.c_str());
    while (dirent* e = ::readdir(d)) {
        std::string nm = e->d_name;
        size_t dot = nm.find_last_of('.');
        if (nm == "." || nm == ".." || dot == std::string::npos || dot == 0) continue;
        std::string ext = lower(nm.substr(dot + 1));
        if (ext == "wav" || ext == "flac" || ext == "mp3") files.push_back(nm);
    }
    ::closedir(d);
    std::sort(files.begin(), files.end());
    if (args.limit_files > 0 && files.size() > args.limit_files) files.resize(args.limit_files);
    WB_REQUIRE(!files.empty(), WB_EINVAL, "No audio files found in %s", args.audio_dir.c_str());

    // file groups in listing order; group g belongs to worker process g % gpus
    std::vector<std::pair<size_t, size_t>> groups;
    for (size_t g0 = 0; g0 < files.size(); g0 += args.file_batch) groups.push_back({g0, std::min(files.size(), g0 + args.file_batch)});
    const size_t world = std::min(args.gpus, groups.size());
    // --intra-op is the reference's "CPU threads per unit of work" knob; here the CPU work of a unit is audio decode
    const Job job{args, files, tk.t, gen, mc, wpath, std::max<size_t>(1, std::min<size_t>(cfg.intra_op, 8))};
    TRACE("%zu files in %zu groups, %zu worker process(es) x %zu in flight", files.size(), groups.size(), world, args.in_flight);
    std::vector<FileResult> results(files.size());
    if (world <= 1) {
        run_rank(job, args.device, groups, 0, 1, results);
    } else {
        // one process per GPU.  The parent has not touched CUDA, so a plain fork() is safe; every child
        // creates its own contexts on its own device and hands its rows back through a file.
        fflush(stdout); fflush(stderr);
        std::vector<pid_t> pids(world);
        std::vector<std::string> row_files(world);
        for (size_t r = 0; r < world; ++r) {
            row_files[r] = args.out_csv + ".rank" + std::to_string(r) + ".rows";
            pid_t pid = ::fork();
            WB_REQUIRE(pid >= 0, WB_EIO, "fork failed for worker %zu", r);
            if (pid == 0) {
                int rc = 1;
                try {
                    int n_dev = 0;
                    CK(wb_device_count(&n_dev));
                    const int dev = (args.device + (int)r) % n_dev;
                    if (args.device + (int)r >= n_dev)
                        fprintf(stderr, "note: worker %zu shares GPU %d (%d visible, --gpus %zu)\n", r, dev, n_dev, args.gpus);
                    run_rank(job, dev, groups, r, world, results);
                    write_rows(row_files[r], results);
                    rc = 0;
                } catch (const std::exception& e) {
                    fprintf(stderr, "Error: [gpu %d] %s\n", args.device + (int)r, e.what());
                }
                fflush(stdout); fflush(stderr);
                ::_exit(rc);
            }
            pids[r] = pid;
        }
        bool ok = true;
        for (size_t r = 0; r < world; ++r) {
            int st = 0;
            if (::waitpid(pids[r], &st, 0) < 0 || !WIFEXITED(st) || WEXITSTATUS(st) != 0) ok = false;
        }
        if (ok) for (size_t r = 0; r < world; ++r) read_rows(row_files[r], results);
        for (size_t r = 0; r < world; ++r) ::unlink(row_files[r].c_str());
        WB_REQUIRE(ok, WB_ESTATE, "a GPU worker process failed");
    }

    TRACE("all groups done");
    struct Row { std::string file; double duration_s, end_to_end_s, rtf; std::string text; };
    std::vector<Row> rows;
    std::vector<double> e2e, load, pre, model, dec, rtfs;
    const std::string txt_dir = parent_of(args.out_csv);
    for (size_t i = 0; i < files.size(); ++i) {                                      // main.rs:1164-1213
        const FileResult& fr = results[i];
        WB_REQUIRE(fr.done, WB_ESTATE, "no result for %s", files[i].c_str());
        const std::string& fnm = files[i];
        const double end_to_end_s = fr.load_s + fr.t.end_to_end_s;
        const double rtf = end_to_end_s / std::max(fr.dur, 1e-9);
        rows.push_back(Row{fnm, std::round(fr.dur * 1000.0) / 1000.0, std::round(end_to_end_s * 10000.0) / 10000.0,
                           std::round(rtf * 1000000.0) / 1000000.0, fr.text});
        load.push_back(fr.load_s); pre.push_back(fr.t.preprocess_s); model.push_back(fr.t.model_only_s); dec.push_back(fr.t.decode_s);
        e2e.push_back(end_to_end_s); rtfs.push_back(rtf);
        if (args.write_txt) {
            size_t dot = fnm.find_last_of('.');
            std::string stem = dot == std::string::npos ? fnm : fnm.substr(0, dot);
            size_t b = fr.text.find_first_not_of(" \t\r\n"), e = fr.text.find_last_not_of(" \t\r\n");
            std::string trimmed = b == std::string::npos ? "" : fr.text.substr(b, e - b + 1);
            write_file(join(txt_dir, stem + ".transcript.txt"), trimmed + "\n");
        }
    }

    {   // CSV (main.rs:1216-1229)
        std::string s = "file,duration_s,end_to_end_s,rtf,text\n";
        for (const Row& r : rows)
            s += csv_field(r.file) + "," + fmt_fixed(r.duration_s, 3) + "," + fmt_fixed(r.end_to_end_s, 4) + "," +
                 fmt_fixed(r.rtf, 6) + "," + csv_field(r.text) + "\n";
        write_file(args.out_csv, s);
    }
    {   // per-file JSON (main.rs:1232): struct order, pretty
        std::string s = "[";
        for (size_t i = 0; i < rows.size(); ++i) {
            const Row& r = rows[i];
            s += std::string(i ? "," : "") + "\n  {\n";
            s += "    \"file\": " + wbjson::escape(r.file) + ",\n";
            s += "    \"duration_s\": " + wbjson::fmt_f64(r.duration_s) + ",\n";
            s += "    \"end_to_end_s\": " + wbjson::fmt_f64(r.end_to_end_s) + ",\n";
            s += "    \"rtf\": " + wbjson::fmt_f64(r.rtf) + ",\n";
            s += "    \"text\": " + wbjson::escape(r.text) + "\n  }";
        }
        s += rows.empty() ? "]" : "\n]";
        write_file(args.out_json, s);
    }
    // summary (main.rs:1235-1259): keys alphabetical at every level (serde_json without preserve_order)
    OrtCfg sorted_dummy = cfg;
    (void)sorted_dummy;
    std::string sum = "{\n";
    sum += "  \"breakdown_s\": {\n";
    sum += "    \"decode_s\": " + stat_json(dec, 4) + ",\n";
    sum += "    \"load_s\": " + stat_json(load, 4) + ",\n";
    sum += "    \"model_only_s\": " + stat_json(model, 4) + ",\n";
    sum += "    \"preprocess_s\": " + stat_json(pre, 4) + "\n  },\n";
    sum += "  \"config_used\": {\n";
    sum += std::string("    \"allow_spinning\": ") + (cfg.allow_spinning ? "true" : "false") + ",\n";
    sum += std::string("    \"cpu_mem_arena\": ") + (cfg.cpu_mem_arena ? "true" : "false") + ",\n";
    sum += "    \"execution_mode\": " + wbjson::escape(cfg.execution_mode) + ",\n";
    sum += "    \"graph_opt\": " + wbjson::escape(cfg.graph_opt) + ",\n";
    sum += "    \"inter_op\": " + std::to_string(cfg.inter_op) + ",\n";
    sum += "    \"intra_op\": " + std::to_string(cfg.intra_op) + ",\n";
    sum += std::string("    \"mem_pattern\": ") + (cfg.mem_pattern ? "true" : "false") + "\n  },\n";
    sum += "  \"language\": " + wbjson::escape(args.language) + ",\n";
    sum += "  \"latency_end_to_end_s\": " + stat_json(e2e, 2) + ",\n";
    sum += "  \"max_new_tokens\": " + std::to_string(args.max_new_tokens) + ",\n";
    sum += "  \"model_id\": " + wbjson::escape(args.model_id) + ",\n";
    sum += "  \"n_files\": " + std::to_string(rows.size()) + ",\n";
    sum += "  \"notes\": {\n";
    sum += "    \"longform\": \"Rust approximation: chunked 30s windows with overlap; greedy decode via decoder_with_past\",\n";
    sum += std::string("    \"token_decode\": ") + (tk.t ? "\"Tokenizer decode (skip_special_tokens=true)\"" : "\"Prints token IDs unless you provide tokenizer.json.\"") + "\n  },\n";
    sum += "  \"onnx_dir\": " + wbjson::escape(args.onnx_dir) + ",\n";
    sum += "  \"rtf_end_to_end\": " + stat_json(rtfs, 2) + ",\n";
    sum += "  \"task\": " + wbjson::escape(args.task) + ",\n";
    sum += std::string("  \"timestamps\": ") + (args.timestamps ? "true" : "false") + ",\n";
    sum += "  \"tokenizer_json\": " + wbjson::escape(tk.path) + "\n}";
    write_file(args.out_summary_json, sum);

    TRACE("outputs written");
    printf("DONE\n");                                                                // main.rs:1261-1268
    printf("Config used:\n%s\n", cfg_json(cfg, 0).c_str());
    printf("Per-file CSV: %s\n", args.out_csv.c_str());
    printf("Per-file JSON: %s\n", args.out_json.c_str());
    printf("Summary JSON: %s\n", args.out_summary_json.c_str());
    printf("End-to-end p95(s): %.6f\n", wb_host_percentile(e2e.data(), (int)e2e.size(), 95.0));
    return 0;
}

}  // namespace

extern "C" int wb_cli_main(int argc, const char* const* argv) {
    try {
        Args a = parse_args(argc, argv);
        return run(a);
    } catch (const UsageError& e) {
        if (std::strlen(e.what()) == 0) { fputs(USAGE, stdout); return 0; }
        fprintf(stderr, "%s\n\n%s", e.what(), USAGE);
        return 2;                                         // clap's usage-error exit status
    } catch (const std::exception& e) {
        fprintf(stderr, "Error: %s\n", e.what());         // anyhow's `Error: ...` + exit status 1
        return 1;
    }
}
