// weights.cpp — weight sources and HBM packing (replaces the three build_session calls,
// /root/reference/src/main.rs:1099-1108: ORT loads the ONNX initializers; here the same tensors
// arrive as a .wb200 blob or from the seeded generator, and are repacked for the kernels).
//
// The generator mirrors whisper-rust-ort_b200/weights.py exactly (same tensor order = tensor id,
// same integer hash, one f32 multiply + one f32 add) so host oracle and device agree bit-for-bit.
#include <cstring>
#include <fstream>
#include <mutex>

#include <sys/stat.h>

#include "ctx.h"
#include "host/json.h"

void onnx_load_dir(const std::string& dir, const wb_model_cfg& c, std::map<std::string, std::vector<float>>& host);   // host/onnx.cpp

namespace {

struct Spec {
    std::string name;
    std::vector<int64_t> shape;
    int kind;       // 0 random, 1 sinusoid
    double std_, off;
};

std::vector<Spec> tensor_specs(const wb_model_cfg& c) {
    const int64_t d = c.d_model, f = c.ffn_dim;
    const double W = 0.02, Bs = 0.02, LNW_ = 0.05, LNB = 0.02;
    std::vector<Spec> s;
    auto lin = [&](const std::string& p, int64_t out, int64_t in, bool bias = true) {
        s.push_back({p + ".weight", {out, in}, 0, W, 0.0});
        if (bias) s.push_back({p + ".bias", {out}, 0, Bs, 0.0});
    };
    auto ln = [&](const std::string& p) {
        s.push_back({p + ".weight", {d}, 0, LNW_, 1.0});
        s.push_back({p + ".bias", {d}, 0, LNB, 0.0});
    };
    auto attn = [&](const std::string& p) {
        lin(p + ".k_proj", d, d, false);
        lin(p + ".v_proj", d, d);
        lin(p + ".q_proj", d, d);
        lin(p + ".out_proj", d, d);
    };
    const std::string e = "model.encoder";
    s.push_back({e + ".conv1.weight", {d, c.n_mels, 3}, 0, W, 0.0});
    s.push_back({e + ".conv1.bias", {d}, 0, Bs, 0.0});
    s.push_back({e + ".conv2.weight", {d, d, 3}, 0, W, 0.0});
    s.push_back({e + ".conv2.bias", {d}, 0, Bs, 0.0});
    s.push_back({e + ".embed_positions.weight", {c.n_audio_ctx, d}, 1, 0.0, 0.0});
    for (int i = 0; i < c.enc_layers; ++i) {
        std::string p = e + ".layers." + std::to_string(i);
        attn(p + ".self_attn");
        ln(p + ".self_attn_layer_norm");
        lin(p + ".fc1", f, d);
        lin(p + ".fc2", d, f);
        ln(p + ".final_layer_norm");
    }
    ln(e + ".layer_norm");
    const std::string dd = "model.decoder";
    s.push_back({dd + ".embed_tokens.weight", {c.vocab, d}, 0, W, 0.0});
    s.push_back({dd + ".embed_positions.weight", {c.n_text_ctx, d}, 0, W, 0.0});
    for (int i = 0; i < c.dec_layers; ++i) {
        std::string p = dd + ".layers." + std::to_string(i);
        attn(p + ".self_attn");
        ln(p + ".self_attn_layer_norm");
        attn(p + ".encoder_attn");
        ln(p + ".encoder_attn_layer_norm");
        lin(p + ".fc1", f, d);
        lin(p + ".fc2", d, f);
        ln(p + ".final_layer_norm");
    }
    ln(dd + ".layer_norm");
    return s;
}

int64_t numel(const Spec& s) {
    int64_t n = 1;
    for (auto v : s.shape) n *= v;
    return n;
}

void generate_tensor(const Spec& sp, int tid, uint64_t seed, std::vector<float>& out) {
    const int64_t n = numel(sp);
    out.resize((size_t)n);
    if (sp.kind == 1) {      // HF `sinusoids`, evaluated in f64 and rounded once
        const int64_t len = sp.shape[0], ch = sp.shape[1], half = ch / 2;
        const double inc = std::log(10000.0) / (double)(half - 1);
        for (int64_t j = 0; j < half; ++j) {
            const double inv = std::exp(-inc * (double)j);
            for (int64_t t = 0; t < len; ++t) {
                const double a = (double)t * inv;
                out[(size_t)(t * ch + j)] = (float)std::sin(a);
                out[(size_t)(t * ch + half + j)] = (float)std::cos(a);
            }
        }
        return;
    }
    const uint64_t GOLDEN = 0x9E3779B97F4A7C15ull, TM = 0xD1B54A32D192ED03ull;
    const uint64_t M1 = 0xBF58476D1CE4E5B9ull, M2 = 0x94D049BB133111EBull;
    const double IH4_SD = std::sqrt(4.0 * (65536.0 * 65536.0 - 1.0) / 12.0);
    const float scale = (float)(sp.std_ / IH4_SD);
    const float off = (float)sp.off;
    const uint64_t key = seed * GOLDEN + (uint64_t)(tid + 1) * TM;
    for (int64_t i = 0; i < n; ++i) {
        uint64_t z = key + (uint64_t)(i + 1) * GOLDEN;
        z = (z ^ (z >> 30)) * M1;
        z = (z ^ (z >> 27)) * M2;
        z = z ^ (z >> 31);
        int32_t s = (int32_t)((z & 0xFFFF) + ((z >> 16) & 0xFFFF) + ((z >> 32) & 0xFFFF) + ((z >> 48) & 0xFFFF)) - 131070;
        volatile float v = (float)s * scale;        // volatile: forbid fma contraction with the add
        out[(size_t)i] = sp.off != 0.0 ? v + off : v;
    }
}

void load_blob(const char* path, const wb_model_cfg& cfg, const std::vector<Spec>& specs,
               std::map<std::string, std::vector<float>>& host) {
    std::ifstream f(path, std::ios::binary);
    WB_REQUIRE(f.good(), WB_EIO, "cannot open weights blob: %s", path);
    char magic[8];
    uint64_t jlen = 0;
    f.read(magic, 8);
    f.read(reinterpret_cast<char*>(&jlen), 8);
    WB_REQUIRE(f.good() && std::memcmp(magic, "WB200W01", 8) == 0 && jlen < (1u << 26), WB_EIO,
               "%s is not a .wb200 weight blob", path);
    std::string js((size_t)jlen, '\0');
    f.read(&js[0], (std::streamsize)jlen);
    WB_REQUIRE(f.good(), WB_EIO, "truncated blob header: %s", path);
    wbjson::Value root = wbjson::parse(js);
    const wbjson::Value& c = root["cfg"];
    auto chk = [&](const char* k, int v) {
        WB_REQUIRE((int)c[k].num() == v, WB_EINVAL, "blob cfg.%s=%d does not match ctx cfg %d", k, (int)c[k].num(), v);
    };
    chk("n_mels", cfg.n_mels); chk("d_model", cfg.d_model); chk("n_heads", cfg.n_heads);
    chk("ffn_dim", cfg.ffn_dim); chk("enc_layers", cfg.enc_layers); chk("dec_layers", cfg.dec_layers);
    chk("vocab", cfg.vocab); chk("n_audio_ctx", cfg.n_audio_ctx); chk("n_text_ctx", cfg.n_text_ctx);
    size_t payload = (16 + (size_t)jlen + 63) / 64 * 64;
    std::map<std::string, std::pair<uint64_t, uint64_t>> idx;
    for (const auto& t : root["tensors"].arr()) idx[t["name"].str()] = {(uint64_t)t["offset"].num(), (uint64_t)t["nbytes"].num()};
    for (const auto& sp : specs) {
        auto it = idx.find(sp.name);
        WB_REQUIRE(it != idx.end(), WB_EINVAL, "blob misses tensor %s", sp.name.c_str());
        WB_REQUIRE(it->second.second == (uint64_t)numel(sp) * 4, WB_EINVAL, "blob tensor %s has wrong size", sp.name.c_str());
        std::vector<float>& v = host[sp.name];
        v.resize((size_t)numel(sp));
        f.seekg((std::streamoff)(payload + it->second.first));
        f.read(reinterpret_cast<char*>(v.data()), (std::streamsize)it->second.second);
        WB_REQUIRE(f.good(), WB_EIO, "truncated blob payload at %s", sp.name.c_str());
    }
}

struct Uploader {
    wb_ctx* ctx;
    bool bf16;
    float* f32(const std::vector<float>& v) {
        float* p = nullptr;
        CUDA_CHECK(cudaMalloc(&p, v.size() * sizeof(float)));
        ctx->w.store->allocs.push_back(p);
        CUDA_CHECK(cudaMemcpy(p, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice));
        return p;
    }
    void* compute(const std::vector<float>& v) {
        if (!bf16) return f32(v);
        std::vector<__nv_bfloat16> h(v.size());
        for (size_t i = 0; i < v.size(); ++i) h[i] = __float2bfloat16(v[i]);
        void* p = nullptr;
        CUDA_CHECK(cudaMalloc(&p, h.size() * 2));
        ctx->w.store->allocs.push_back(p);
        CUDA_CHECK(cudaMemcpy(p, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
        return p;
    }
};

}  // namespace

WeightStore::~WeightStore() {
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(device);
    for (void* p : allocs) cudaFree(p);
    cudaSetDevice(prev);
}

namespace {
std::mutex g_store_mu;
std::map<std::string, std::weak_ptr<WeightStore>> g_stores;

// identity of an uploaded model: device, precision, architecture, and the source (seed, or path + mtime + size)
std::string store_key(const wb_ctx* ctx, const char* path) {
    const wb_model_cfg& c = ctx->cfg;
    char b[512];
    snprintf(b, sizeof(b), "dev%d|p%d|%d.%d.%d.%d.%d.%d.%d.%d.%d|", ctx->device, c.precision, c.n_mels, c.d_model, c.n_heads, c.ffn_dim,
             c.enc_layers, c.dec_layers, c.vocab, c.n_audio_ctx, c.n_text_ctx);
    std::string k = b;
    struct stat st;
    if (path && path[0]) {
        k += std::string("path:") + path;
        if (::stat(path, &st) == 0) k += "|" + std::to_string((long long)st.st_mtime) + "|" + std::to_string((long long)st.st_size);
    } else {
        k += "seed:" + std::to_string((unsigned long long)c.seed);
    }
    return k;
}
void weights_build(wb_ctx* ctx, const char* path);
}  // namespace

void weights_init(wb_ctx* ctx, const char* path) {
    // creation is serialised: the first context of a (device, source) builds and uploads, the others wait and share
    std::lock_guard<std::mutex> lk(g_store_mu);
    const std::string key = store_key(ctx, path);
    auto it = g_stores.find(key);
    if (it != g_stores.end()) {
        if (std::shared_ptr<WeightStore> st = it->second.lock()) {
            ctx->w = st->view;
            ctx->w.store = st;
            return;
        }
    }
    ctx->w = ModelW{};
    ctx->w.store = std::make_shared<WeightStore>();
    ctx->w.store->device = ctx->device;
    weights_build(ctx, path);
    ctx->w.store->view = ctx->w;
    ctx->w.store->view.store.reset();
    g_stores[key] = ctx->w.store;
}

namespace {
void weights_build(wb_ctx* ctx, const char* path) {
    const wb_model_cfg& c = ctx->cfg;
    auto specs = tensor_specs(c);
    auto& host = ctx->w.store->host;
    struct stat st;
    if (path && path[0] && ::stat(path, &st) == 0 && S_ISDIR(st.st_mode)) {
        onnx_load_dir(path, c, host);                       // optimum export: encoder_model.onnx + decoder_model.onnx
        for (const auto& sp : specs) {
            auto it = host.find(sp.name);
            WB_REQUIRE(it != host.end() && (int64_t)it->second.size() == numel(sp), WB_EINVAL, "ONNX export did not yield tensor %s", sp.name.c_str());
        }
    } else if (path && path[0]) {
        load_blob(path, c, specs, host);
    } else {
        for (size_t i = 0; i < specs.size(); ++i) generate_tensor(specs[i], (int)i, c.seed, host[specs[i].name]);
    }
    Uploader up{ctx, c.precision == WB_PREC_BF16};
    const int d = c.d_model;
    auto H = [&](const std::string& n) -> const std::vector<float>& {
        auto it = host.find(n);
        WB_REQUIRE(it != host.end(), WB_EINVAL, "missing tensor %s", n.c_str());
        return it->second;
    };
    auto lnw = [&](const std::string& p) { return LNW{up.f32(H(p + ".weight")), up.f32(H(p + ".bias"))}; };
    auto linear = [&](const std::string& p, int out, int in, bool bias = true) {
        LinearW l;
        l.w = up.compute(H(p + ".weight"));
        l.b = bias ? up.f32(H(p + ".bias")) : nullptr;
        l.out = out; l.in = in;
        return l;
    };
    // conv [co][ci][k] -> GEMM weight [co][k*C + ci] (A rows are 3 consecutive time-major frames)
    auto conv = [&](const std::string& p, int co, int ci) {
        const auto& w = H(p + ".weight");
        std::vector<float> packed((size_t)co * 3 * ci);
        for (int o = 0; o < co; ++o)
            for (int i = 0; i < ci; ++i)
                for (int k = 0; k < 3; ++k) packed[((size_t)o * 3 + k) * ci + i] = w[((size_t)o * ci + i) * 3 + k];
        LinearW l;
        l.w = up.compute(packed);
        l.b = up.f32(H(p + ".bias"));
        l.out = co; l.in = 3 * ci;
        return l;
    };
    // fused projections: rows of the listed Linear layers stacked; missing bias = zeros
    auto fused = [&](const std::vector<std::string>& ps, const std::vector<bool>& has_bias) {
        std::vector<float> w, b;
        for (size_t i = 0; i < ps.size(); ++i) {
            const auto& wi = H(ps[i] + ".weight");
            w.insert(w.end(), wi.begin(), wi.end());
            if (has_bias[i]) {
                const auto& bi = H(ps[i] + ".bias");
                b.insert(b.end(), bi.begin(), bi.end());
            } else {
                b.insert(b.end(), (size_t)d, 0.0f);
            }
        }
        LinearW l;
        l.w = up.compute(w);
        l.b = up.f32(b);
        l.out = (int)ps.size() * d; l.in = d;
        return l;
    };
    ModelW& m = ctx->w;
    const std::string e = "model.encoder";
    m.conv1 = conv(e + ".conv1", d, c.n_mels);
    m.conv2 = conv(e + ".conv2", d, d);
    m.enc_pos = up.f32(H(e + ".embed_positions.weight"));
    m.enc.resize(c.enc_layers);
    for (int i = 0; i < c.enc_layers; ++i) {
        std::string p = e + ".layers." + std::to_string(i);
        EncLayerW& L = m.enc[i];
        L.ln1 = lnw(p + ".self_attn_layer_norm");
        L.qkv = fused({p + ".self_attn.q_proj", p + ".self_attn.k_proj", p + ".self_attn.v_proj"}, {true, false, true});
        L.o = linear(p + ".self_attn.out_proj", d, d);
        L.ln2 = lnw(p + ".final_layer_norm");
        L.fc1 = linear(p + ".fc1", c.ffn_dim, d);
        L.fc2 = linear(p + ".fc2", d, c.ffn_dim);
    }
    m.enc_ln = lnw(e + ".layer_norm");
    const std::string dd = "model.decoder";
    m.embed = up.compute(H(dd + ".embed_tokens.weight"));
    m.dec_pos = up.f32(H(dd + ".embed_positions.weight"));
    m.dec.resize(c.dec_layers);
    for (int i = 0; i < c.dec_layers; ++i) {
        std::string p = dd + ".layers." + std::to_string(i);
        DecLayerW& L = m.dec[i];
        L.ln1 = lnw(p + ".self_attn_layer_norm");
        L.qkv = fused({p + ".self_attn.q_proj", p + ".self_attn.k_proj", p + ".self_attn.v_proj"}, {true, false, true});
        L.o = linear(p + ".self_attn.out_proj", d, d);
        L.ln2 = lnw(p + ".encoder_attn_layer_norm");
        L.cq = linear(p + ".encoder_attn.q_proj", d, d);
        L.ckv = fused({p + ".encoder_attn.k_proj", p + ".encoder_attn.v_proj"}, {false, true});
        L.co = linear(p + ".encoder_attn.out_proj", d, d);
        L.ln3 = lnw(p + ".final_layer_norm");
        L.fc1 = linear(p + ".fc1", c.ffn_dim, d);
        L.fc2 = linear(p + ".fc2", d, c.ffn_dim);
    }
    m.dec_ln = lnw(dd + ".layer_norm");
}

}  // namespace

void weights_free(wb_ctx* ctx) {
    ctx->w = ModelW{};          // the store frees the device copy when its last context goes
}
