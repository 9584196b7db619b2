// api.cpp — the extern "C" surface of include/whisper_b200.h over the CUDA path.
// No exception crosses the boundary: every entry point converts to a status code + message.
#include <cmath>
#include <cstring>
#include <mutex>

#include <chrono>
#include "ctx.h"

static thread_local std::string g_err;
void wb_set_error(const std::string& msg) { g_err = msg; }

#define WB_TRY try {
#define WB_CATCH                                                             \
    }                                                                        \
    catch (const WbError& e) { wb_set_error(e.what()); return e.code; }      \
    catch (const std::exception& e) { wb_set_error(e.what()); return WB_EINVAL; } \
    return WB_OK;

namespace {

// Every entry point may be called from a different host thread than the one that created the ctx; CUDA's
// current device is per thread, so bind it to the ctx's device first (streams/events are device-bound).
void require_ctx(const wb_ctx* c) {
    WB_REQUIRE(c != nullptr, WB_EINVAL, "null wb_ctx");
    CUDA_CHECK(cudaSetDevice(c->device));
}

// Stage file list + chunk table on the device (wb_upload_pcm).
void upload_pcm(wb_ctx* ctx, const float* pcm, const int64_t* offsets, int n_files, int64_t chunk_len,
                int64_t step, int* n_chunks_out) {
    WB_REQUIRE(pcm && offsets && n_files > 0, WB_EINVAL, "wb_upload_pcm: bad arguments");
    const int NM = ctx->cfg.n_mels;
    WB_REQUIRE(NM == 80 || NM == 128, WB_EINVAL, "log-mel kernel has 80 bins (the reference's frontend) or 128 (large-v3), not %d", NM);
    if (chunk_len <= 0) chunk_len = WB_CHUNK_SAMPLES;
    if (step <= 0) step = 400000;
    WB_REQUIRE(offsets[0] == 0, WB_EINVAL, "offsets[0] must be 0");
    // build and validate into locals; the context's MelState changes only once everything checked out (a rejected
    // upload must not leave metadata that disagrees with the device buffers of the previous one)
    std::vector<int64_t> h_frame_off(n_files + 1, 0), h_chunk_pos;
    std::vector<int> h_tile_off(n_files + 1, 0);
    std::vector<MelChunk> h_chunks;
    for (int i = 0; i < n_files; ++i) {
        const int64_t n = offsets[i + 1] - offsets[i];
        WB_REQUIRE(n > 0, WB_EINVAL, "Empty audio (file %d)", i);                      // main.rs:414-416
        const int64_t nf = mel_n_frames(n);
        h_frame_off[i + 1] = h_frame_off[i] + nf;
        h_tile_off[i + 1] = h_tile_off[i] + (int)ceil_div64(nf, WB_MEL_FPT);
        int64_t pos = 0;                                                               // main.rs:875-882
        while (pos < n) {
            const int64_t end = pos + chunk_len < n ? pos + chunk_len : n;
            h_chunks.push_back(MelChunk{i, (int)(pos / 160)});
            h_chunk_pos.push_back(pos);
            if (end == n) break;
            pos += step;
        }
    }
    WB_REQUIRE((int)h_chunks.size() <= ctx->cfg.max_chunks, WB_ECAP, "%d chunks exceed max_chunks %d", (int)h_chunks.size(), ctx->cfg.max_chunks);
    MelState& s = ctx->mel;
    s.raw_valid = false;
    s.n_files = 0;                                   // stays 0 (= nothing resident) if an allocation or copy below fails
    s.n_chunks = 0;
    s.h_file_off.assign(offsets, offsets + n_files + 1);
    s.h_frame_off.swap(h_frame_off);
    s.h_tile_off.swap(h_tile_off);
    s.h_chunks.swap(h_chunks);
    s.h_chunk_pos.swap(h_chunk_pos);
    const int n_chunks = (int)s.h_chunks.size();
    s.total_frames = s.h_frame_off[n_files];
    s.total_tiles = s.h_tile_off[n_files];
    const int64_t total = offsets[n_files] - offsets[0];
    s.pcm.reserve((size_t)total);
    s.file_off.reserve(n_files + 1);
    s.frame_off.reserve(n_files + 1);
    s.tile_off.reserve(n_files + 1);
    s.fmax.reserve(n_files);
    s.raw.reserve((size_t)s.total_frames * NM);
    s.chunks.reserve(n_chunks);
    cudaStream_t st = ctx->stream;
    CudaEvent e0, e1;
    CUDA_CHECK(cudaEventRecord(e0.e, st));
    CUDA_CHECK(cudaMemcpyAsync(s.pcm.p, pcm, sizeof(float) * (size_t)total, cudaMemcpyHostToDevice, st));
    CUDA_CHECK(cudaMemcpyAsync(s.file_off.p, s.h_file_off.data(), sizeof(int64_t) * (n_files + 1), cudaMemcpyHostToDevice, st));
    CUDA_CHECK(cudaMemcpyAsync(s.frame_off.p, s.h_frame_off.data(), sizeof(int64_t) * (n_files + 1), cudaMemcpyHostToDevice, st));
    CUDA_CHECK(cudaMemcpyAsync(s.tile_off.p, s.h_tile_off.data(), sizeof(int) * (n_files + 1), cudaMemcpyHostToDevice, st));
    CUDA_CHECK(cudaMemcpyAsync(s.chunks.p, s.h_chunks.data(), sizeof(MelChunk) * n_chunks, cudaMemcpyHostToDevice, st));
    s.n_files_staged = n_files;
    mel_build_tiles(ctx);
    CUDA_CHECK(cudaEventRecord(e1.e, st));
    CUDA_CHECK(wb_stream_sync(st));
    CUDA_CHECK(cudaEventElapsedTime(&ctx->timing.h2d_ms, e0.e, e1.e));
    s.n_files = n_files;                             // committed: the device buffers now match the metadata
    s.n_chunks = n_chunks;
    if (n_chunks_out) *n_chunks_out = s.n_chunks;
}

void run_log_mel(wb_ctx* ctx, bool sync) {
    MelState& s = ctx->mel;
    WB_REQUIRE(s.n_files > 0, WB_ESTATE, "wb_run_log_mel before wb_upload_pcm");
    ctx->timing.mel_launches = 0;
    CUDA_CHECK(cudaEventRecord(ctx->mel_e0.e, ctx->stream));
    mel_launch_raw(ctx);
    mel_launch_chunks(ctx, 0, s.n_chunks, ctx->enc.mel_tm.p);
    CUDA_CHECK(cudaEventRecord(ctx->mel_e1.e, ctx->stream));
    s.raw_valid = true;
    ctx->mel_t_pending = true;
    if (sync) timing_flush(ctx);
}

DecodeParams make_params(int B, const int64_t* prompt, int prompt_len, int max_new, int64_t eot,
                         const int64_t* sup, int ns, const int64_t* bsup, int nb, const int64_t* forced,
                         bool want_logits) {
    WB_REQUIRE(prompt && prompt_len > 0, WB_EINVAL, "empty prompt");
    WB_REQUIRE((ns == 0 || sup) && (nb == 0 || bsup), WB_EINVAL, "null suppress list");
    DecodeParams p{};
    p.B = B; p.prompt = prompt; p.prompt_len = prompt_len; p.max_new = max_new; p.eot = (int)eot;
    p.suppress = sup; p.n_suppress = ns; p.begin_suppress = bsup; p.n_begin_suppress = nb;
    p.forced = forced; p.want_logits = want_logits;
    return p;
}

void transcribe_resident(wb_ctx* ctx, const DecodeParams& proto, int64_t* tokens_out, int32_t* lens_out, int cap_chunks) {
    MelState& s = ctx->mel;
    WB_REQUIRE(s.n_chunks <= cap_chunks, WB_ECAP, "output capacity %d < %d chunks", cap_chunks, s.n_chunks);
    run_log_mel(ctx, false);             // the encoder is enqueued right behind it; mel_ms is read back after the decode
    const int stride = proto.prompt_len + (proto.max_new < 1 ? 1 : proto.max_new);
    const size_t chunk_elems = (size_t)(WB_N_FRAMES + 2) * ctx->cfg.n_mels * ctx->esz();
    float enc_ms = 0, ckv_ms = 0, dec_ms = 0;
    int dl = 0, el = 0, ds = 0;
    for (int c0 = 0; c0 < s.n_chunks; c0 += ctx->cfg.max_batch) {
        const int B = s.n_chunks - c0 < ctx->cfg.max_batch ? s.n_chunks - c0 : ctx->cfg.max_batch;
        encoder_forward(ctx, (const char*)ctx->enc.mel_tm.p + (size_t)c0 * chunk_elems, B);
        DecodeParams p = proto;
        p.B = B;
        decoder_run(ctx, p);
        decoder_fetch(ctx, p, tokens_out + (size_t)c0 * stride, lens_out + c0, nullptr);
        enc_ms += ctx->timing.encoder_ms; ckv_ms += ctx->timing.cross_kv_ms; dec_ms += ctx->timing.decode_ms;
        dl += ctx->timing.decode_launches; el += ctx->timing.encoder_launches; ds += ctx->timing.decode_steps;
    }
    ctx->timing.encoder_ms = enc_ms; ctx->timing.cross_kv_ms = ckv_ms; ctx->timing.decode_ms = dec_ms;
    ctx->timing.decode_launches = dl; ctx->timing.encoder_launches = el; ctx->timing.decode_steps = ds;
}

}  // namespace

void timing_flush(wb_ctx* ctx) {
    if (ctx->mel_t_pending) {
        CUDA_CHECK(cudaEventSynchronize(ctx->mel_e1.e));
        CUDA_CHECK(cudaEventElapsedTime(&ctx->timing.mel_ms, ctx->mel_e0.e, ctx->mel_e1.e));
        ctx->mel_t_pending = false;
    }
    if (ctx->enc_t_pending) {
        CUDA_CHECK(cudaEventSynchronize(ctx->enc_e2.e));
        CUDA_CHECK(cudaEventElapsedTime(&ctx->timing.encoder_ms, ctx->enc_e0.e, ctx->enc_e1.e));
        CUDA_CHECK(cudaEventElapsedTime(&ctx->timing.cross_kv_ms, ctx->enc_e1.e, ctx->enc_e2.e));
        ctx->enc_t_pending = false;
    }
}

extern "C" {

const char* wb_last_error(void) { return g_err.c_str(); }

int wb_host_alloc_pinned(int device, size_t bytes, void** out) {
    WB_TRY
    WB_REQUIRE(out != nullptr && bytes > 0, WB_EINVAL, "wb_host_alloc_pinned: bad arguments");
    *out = nullptr;
    CUDA_CHECK(cudaSetDevice(device));
    CUDA_CHECK(cudaHostAlloc(out, bytes, cudaHostAllocPortable));
    WB_CATCH
}

void wb_host_free_pinned(void* p) { if (p) cudaFreeHost(p); }

int wb_device_count(int* count_out) {
    WB_TRY
    WB_REQUIRE(count_out != nullptr, WB_EINVAL, "count_out is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        *count_out = 0;
        WB_THROW(WB_ECUDA, "no CUDA device: %s", e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    *count_out = n;
    WB_CATCH
}

int wb_default_cfg(wb_model_cfg* cfg, const char* name) {
    WB_TRY
    WB_REQUIRE(cfg && name, WB_EINVAL, "wb_default_cfg: null argument");
    std::memset(cfg, 0, sizeof(*cfg));
    const std::string n(name);
    if (n == "base") *cfg = wb_model_cfg{80, 512, 8, 2048, 6, 6, 51865, 1500, 448, WB_PREC_FP32, 32, 64, 0};
    else if (n == "large-v3") *cfg = wb_model_cfg{128, 1280, 20, 5120, 32, 32, 51866, 1500, 448, WB_PREC_FP32, 16, 16, 0};
    else if (n == "toy") *cfg = wb_model_cfg{80, 128, 2, 256, 2, 2, 1031, 1500, 448, WB_PREC_FP32, 4, 8, 0};
    else WB_THROW(WB_EINVAL, "unknown model name '%s' (base | large-v3 | toy)", name);
    WB_CATCH
}

int wb_create(wb_ctx** out, int device, const wb_model_cfg* cfg, const char* weights_path) {
    wb_ctx* ctx = nullptr;
    try {
        WB_REQUIRE(out && cfg, WB_EINVAL, "wb_create: null argument");
        *out = nullptr;
        int n_dev = 0;
        cudaError_t e = cudaGetDeviceCount(&n_dev);
        WB_REQUIRE(e == cudaSuccess && n_dev > 0, WB_ECUDA,
                   "no CUDA device available (%s): libwhisper_b200 has no CPU fallback", cudaGetErrorString(e));
        WB_REQUIRE(device >= 0 && device < n_dev, WB_EINVAL, "device %d out of range (%d devices)", device, n_dev);
        CUDA_CHECK(cudaSetDevice(device));
        cudaDeviceProp prop{};
        CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
        WB_REQUIRE(prop.major == 10, WB_ECUDA, "device %s is sm_%d%d; this library is built for sm_100a (B200) only",
                   prop.name, prop.major, prop.minor);
        WB_REQUIRE(cfg->precision == WB_PREC_FP32 || cfg->precision == WB_PREC_BF16, WB_EINVAL, "bad precision");
        WB_REQUIRE(cfg->max_batch >= 1 && cfg->max_chunks >= cfg->max_batch, WB_EINVAL, "need max_chunks >= max_batch >= 1");
        ctx = new wb_ctx();
        ctx->cfg = *cfg;
        ctx->device = device;
        ctx->sm_count = prop.multiProcessorCount;
        CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        // WB_TRACE_CREATE=1: where context creation spends its time (stderr)
        const bool tr = getenv("WB_TRACE_CREATE") != nullptr;
        auto t_last = std::chrono::steady_clock::now();
        auto lap = [&](const char* what) {
            if (!tr) return;
            auto now = std::chrono::steady_clock::now();
            fprintf(stderr, "[wb_create dev %d] %-14s %.3f s\n", device, what, std::chrono::duration<double>(now - t_last).count());
            t_last = now;
        };
        lap("cuda init");
        // >48 KB dynamic shared memory opt-ins are per device: set them for THIS context's device
        mel_set_attrs();
        gemm_tc_set_attrs();
        attn_tc_set_attrs();
        mel_build_tables(ctx->mel_tables, ctx->cfg.n_mels);
        CUDA_CHECK(cudaMalloc(&ctx->mel_tables_dev, sizeof(MelTables)));
        CUDA_CHECK(cudaMemcpy(ctx->mel_tables_dev, &ctx->mel_tables, sizeof(MelTables), cudaMemcpyHostToDevice));
        lap("mel tables");
        weights_init(ctx, weights_path);
        lap("weights");
        encoder_alloc(ctx);
        lap("encoder alloc");
        decoder_alloc(ctx);
        lap("decoder alloc");
        const char* dbg = getenv("WB_DEBUG");
        ctx->debug = dbg && dbg[0] == '1';
        // not cudaDeviceSynchronize(): another context of this process may be capturing its decode graph on this
        // device right now, and a device-wide sync is illegal while any stream captures
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        CUDA_CHECK(cudaStreamSynchronize(cudaStreamLegacy));
        *out = ctx;
        return WB_OK;
    } catch (const WbError& e) {
        wb_set_error(e.what());
        if (ctx) wb_destroy(ctx);
        return e.code;
    } catch (const std::exception& e) {
        wb_set_error(e.what());
        if (ctx) wb_destroy(ctx);
        return WB_EINVAL;
    }
}

void wb_destroy(wb_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);          // own work only (see wb_create)
    dec_cluster_free(ctx);
    weights_free(ctx);
    for (auto& g : ctx->dec.graphs) cudaGraphExecDestroy(g.exec);
    if (ctx->dec.unfinished_host) cudaFreeHost(ctx->dec.unfinished_host);
    if (ctx->dec.stage_host) cudaFreeHost(ctx->dec.stage_host);
    vocab_tc_free(ctx);
    if (ctx->mel_tables_dev) cudaFree(ctx->mel_tables_dev);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int wb_get_cfg(const wb_ctx* ctx, wb_model_cfg* out) {
    WB_TRY
    require_ctx(ctx);
    WB_REQUIRE(out, WB_EINVAL, "null out");
    *out = ctx->cfg;
    WB_CATCH
}

int wb_get_timing(const wb_ctx* ctx, wb_timing* out) {
    WB_TRY
    require_ctx(ctx);
    WB_REQUIRE(out, WB_EINVAL, "null out");
    timing_flush(const_cast<wb_ctx*>(ctx));        // waits for the stages whose events are still outstanding
    *out = ctx->timing;
    WB_CATCH
}

int wb_mark(wb_ctx* ctx, int slot) {
    WB_TRY
    require_ctx(ctx);
    WB_REQUIRE(slot >= 0 && slot < 8, WB_EINVAL, "slot out of range");
    CUDA_CHECK(cudaEventRecord(ctx->marks[slot].e, ctx->stream));
    WB_CATCH
}

int wb_elapsed_ms(wb_ctx* ctx, int a, int b, float* ms_out) {
    WB_TRY
    require_ctx(ctx);
    WB_REQUIRE(a >= 0 && a < 8 && b >= 0 && b < 8 && ms_out, WB_EINVAL, "bad argument");
    CUDA_CHECK(cudaEventSynchronize(ctx->marks[b].e));
    CUDA_CHECK(cudaEventElapsedTime(ms_out, ctx->marks[a].e, ctx->marks[b].e));
    WB_CATCH
}

int wb_bench_kernel(wb_ctx* ctx, const char* kernel, int B, int iters, float* avg_ms_out, double* bytes_out) {
    WB_TRY
    require_ctx(ctx);
    WB_REQUIRE(kernel && avg_ms_out && bytes_out && iters > 0, WB_EINVAL, "bad argument");
    if (std::string(kernel) == "logmel") {
        WB_REQUIRE(ctx->mel.n_files > 0, WB_ESTATE, "no PCM resident");
        CudaEvent e0, e1;
        mel_launch_raw(ctx);
        CUDA_CHECK(cudaEventRecord(e0.e, ctx->stream));
        for (int i = 0; i < iters; ++i) mel_launch_raw(ctx);
        CUDA_CHECK(cudaEventRecord(e1.e, ctx->stream));
        CUDA_CHECK(cudaEventSynchronize(e1.e));
        float ms = 0;
        CUDA_CHECK(cudaEventElapsedTime(&ms, e0.e, e1.e));
        *avg_ms_out = ms / iters;
        *bytes_out = 4.0 * (double)(ctx->mel.h_file_off[ctx->mel.n_files]) + 4.0 * (double)ctx->cfg.n_mels * (double)ctx->mel.total_frames;
    } else {
        decoder_bench(ctx, kernel, B, iters, avg_ms_out, bytes_out);
    }
    WB_CATCH
}

int wb_selftest_gemm(wb_ctx* ctx, int M, int N, int K, int lda, int batch, int flags, float* max_diff_out, float* max_abs_out) {
    WB_TRY
    require_ctx(ctx);
    WB_REQUIRE(M > 0 && N > 0 && K > 0 && lda > 0 && batch > 0 && max_diff_out && max_abs_out, WB_EINVAL, "bad argument");
    const bool f32out = flags & 1;
    // A rows may overlap (lda < K), as in the conv stem: allocate what the last row touches
    const size_t a_batch = (size_t)(M - 1) * lda + K + 8, a_elems = a_batch * batch;
    std::vector<__nv_bfloat16> hA(a_elems), hW((size_t)N * K);
    std::vector<float> hb(N), hres((size_t)batch * M * N);
    uint32_t st = 12345u;
    auto rnd = [&]() { st = st * 1664525u + 1013904223u; return ((st >> 8) & 0xFFFF) / 65536.0f - 0.5f; };
    for (auto& v : hA) v = __float2bfloat16(rnd());
    for (auto& v : hW) v = __float2bfloat16(rnd() * 0.25f);
    for (auto& v : hb) v = rnd();
    for (auto& v : hres) v = rnd();
    DevBuf<__nv_bfloat16> dA, dW;
    DevBuf<float> db, dres0, dres1;
    DevBuf<unsigned char> c0, c1;
    const size_t out_elems = (size_t)batch * M * N, esz = f32out ? 4 : 2;
    dA.reserve(a_elems); dW.reserve(hW.size()); db.reserve(N); dres0.reserve(out_elems); dres1.reserve(out_elems);
    c0.reserve(out_elems * esz); c1.reserve(out_elems * esz);
    CUDA_CHECK(cudaMemcpy(dA.p, hA.data(), a_elems * 2, cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMemcpy(dW.p, hW.data(), hW.size() * 2, cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMemcpy(db.p, hb.data(), N * 4, cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMemcpy(dres0.p, hres.data(), out_elems * 4, cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMemcpy(dres1.p, hres.data(), out_elems * 4, cudaMemcpyHostToDevice));
    GemmArgs g;
    g.A = dA.p; g.B = dW.p; g.ta = g.tb = WB_BF16; g.tc = f32out ? WB_F32 : WB_BF16;
    g.M = M; g.N = N; g.K = K; g.lda = lda; g.ldb = K; g.ldc = N; g.batch = batch; g.inner = 1;
    g.sAo = (long long)a_batch; g.sCo = (long long)M * N; g.bias = db.p; g.act = 1;
    WB_REQUIRE(gemm_tc_eligible(g), WB_EINVAL, "shape not eligible for the tcgen05 kernel");
    g.C = c0.p; g.residual = dres0.p;
    gemm_tc(ctx, g);
    g.C = c1.p; g.residual = dres1.p;
    gemm_simt(ctx, g);
    CUDA_CHECK(wb_stream_sync(ctx->stream));
    std::vector<unsigned char> h0(out_elems * esz), h1(out_elems * esz);
    CUDA_CHECK(cudaMemcpy(h0.data(), c0.p, h0.size(), cudaMemcpyDeviceToHost));
    CUDA_CHECK(cudaMemcpy(h1.data(), c1.p, h1.size(), cudaMemcpyDeviceToHost));
    float md = 0.f, ma = 0.f;
    for (size_t i = 0; i < out_elems; ++i) {
        float x0, x1;
        if (f32out) { x0 = reinterpret_cast<float*>(h0.data())[i]; x1 = reinterpret_cast<float*>(h1.data())[i]; }
        else { x0 = __bfloat162float(reinterpret_cast<__nv_bfloat16*>(h0.data())[i]); x1 = __bfloat162float(reinterpret_cast<__nv_bfloat16*>(h1.data())[i]); }
        float d = std::fabs(x0 - x1);
        if (!(d <= md)) md = d;            // NaN-propagating max
        if (std::fabs(x1) > ma) ma = std::fabs(x1);
    }
    *max_diff_out = md; *max_abs_out = ma;
    WB_CATCH
}

int wb_selftest_attn(wb_ctx* ctx, int B, float* max_diff_out, float* max_abs_out) {
    WB_TRY
    require_ctx(ctx);
    const wb_model_cfg& c = ctx->cfg;
    WB_REQUIRE(c.precision == WB_PREC_BF16, WB_EINVAL, "wb_selftest_attn needs the bf16 build");
    WB_REQUIRE(B >= 1 && B <= c.max_batch && max_diff_out && max_abs_out, WB_EINVAL, "bad argument");
    const int T = c.n_audio_ctx, d = c.d_model;
    const size_t n_in = (size_t)B * T * 3 * d, n_out = (size_t)B * T * d;
    std::vector<__nv_bfloat16> h(n_in);
    uint32_t st = 777u;
    for (auto& v : h) { st = st * 1664525u + 1013904223u; v = __float2bfloat16((((st >> 8) & 0xFFFF) / 65536.0f - 0.5f) * 4.0f); }
    CUDA_CHECK(cudaMemcpy(ctx->enc.qkv.p, h.data(), n_in * 2, cudaMemcpyHostToDevice));
    DevBuf<__nv_bfloat16> o0, o1;
    o0.reserve(n_out); o1.reserve(n_out);
    CUDA_CHECK(cudaMemset(o0.p, 0xFF, n_out * 2));
    attn_tc(ctx, ctx->enc.qkv.p, o0.p, B, T, d, c.n_heads);
    attention_simt(ctx, ctx->enc.qkv.p, o1.p, B);
    CUDA_CHECK(wb_stream_sync(ctx->stream));
    std::vector<__nv_bfloat16> h0(n_out), h1(n_out);
    CUDA_CHECK(cudaMemcpy(h0.data(), o0.p, n_out * 2, cudaMemcpyDeviceToHost));
    CUDA_CHECK(cudaMemcpy(h1.data(), o1.p, n_out * 2, cudaMemcpyDeviceToHost));
    float md = 0.f, ma = 0.f;
    for (size_t i = 0; i < n_out; ++i) {
        float x0 = __bfloat162float(h0[i]), x1 = __bfloat162float(h1[i]);
        float dd = std::fabs(x0 - x1);
        if (!(dd <= md)) md = dd;
        if (std::fabs(x1) > ma) ma = std::fabs(x1);
    }
    *max_diff_out = md; *max_abs_out = ma;
    WB_CATCH
}

int wb_set_debug(wb_ctx* ctx, int on) {
    WB_TRY
    require_ctx(ctx);
    ctx->debug = on != 0;
    WB_CATCH
}

int wb_set_load_hint(wb_ctx* ctx, int batches_in_flight) {
    WB_TRY
    require_ctx(ctx);
    WB_REQUIRE(batches_in_flight >= 0, WB_EINVAL, "wb_set_load_hint: negative hint");
    ctx->dec.load_hint = batches_in_flight;
    WB_CATCH
}

int wb_get_tensor(wb_ctx* ctx, const char* name, float* out, int64_t n) {
    WB_TRY
    require_ctx(ctx);
    WB_REQUIRE(ctx->w.store != nullptr, WB_ESTATE, "context has no weights");
    auto it = ctx->w.store->host.find(name ? name : "");
    WB_REQUIRE(it != ctx->w.store->host.end(), WB_EINVAL, "unknown tensor '%s'", name ? name : "");
    WB_REQUIRE((int64_t)it->second.size() == n, WB_EINVAL, "tensor '%s' has %zu elements, caller asked for %lld",
               name, it->second.size(), (long long)n);
    std::memcpy(out, it->second.data(), sizeof(float) * (size_t)n);
    WB_CATCH
}

int wb_upload_pcm(wb_ctx* ctx, const float* pcm, const int64_t* offsets, int n_files, int64_t chunk_len,
                  int64_t step, int* n_chunks_out) {
    WB_TRY
    require_ctx(ctx);
    upload_pcm(ctx, pcm, offsets, n_files, chunk_len, step, n_chunks_out);
    WB_CATCH
}

int wb_run_log_mel(wb_ctx* ctx) {
    WB_TRY
    require_ctx(ctx);
    run_log_mel(ctx, true);
    WB_CATCH
}

int wb_log_mel(wb_ctx* ctx, const float* pcm, const int64_t* offsets, int n_files, int64_t chunk_len,
               int64_t step, float* mel_out, int64_t* n_frames_out, int* n_chunks_out) {
    WB_TRY
    require_ctx(ctx);
    upload_pcm(ctx, pcm, offsets, n_files, chunk_len, step, n_chunks_out);
    run_log_mel(ctx, false);              // mel_ms is read back by the next wait (wb_get_timing, the decode, the copy below)
    MelState& s = ctx->mel;
    if (n_frames_out)
        for (int i = 0; i < n_files; ++i) n_frames_out[i] = s.h_frame_off[i + 1] - s.h_frame_off[i];
    if (mel_out) {
        const int NM = ctx->cfg.n_mels;
        s.export_buf.reserve((size_t)s.total_frames * NM);
        for (int i = 0; i < n_files; ++i) {
            const int64_t nf = s.h_frame_off[i + 1] - s.h_frame_off[i];
            mel_launch_export(ctx, i, 0, nf, s.export_buf.p + s.h_frame_off[i] * NM);
        }
        CudaEvent e0, e1;
        CUDA_CHECK(cudaEventRecord(e0.e, ctx->stream));
        CUDA_CHECK(cudaMemcpyAsync(mel_out, s.export_buf.p, sizeof(float) * (size_t)s.total_frames * NM, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_CHECK(cudaEventRecord(e1.e, ctx->stream));
        CUDA_CHECK(wb_stream_sync(ctx->stream));
        CUDA_CHECK(cudaEventElapsedTime(&ctx->timing.d2h_ms, e0.e, e1.e));
    }
    WB_CATCH
}

int wb_get_chunks(wb_ctx* ctx, int32_t* file_idx, int64_t* sample_pos, int cap) {
    WB_TRY
    require_ctx(ctx);
    MelState& s = ctx->mel;
    WB_REQUIRE(cap >= s.n_chunks, WB_ECAP, "capacity %d < %d chunks", cap, s.n_chunks);
    for (int i = 0; i < s.n_chunks; ++i) {
        if (file_idx) file_idx[i] = s.h_chunks[i].file;
        if (sample_pos) sample_pos[i] = s.h_chunk_pos[i];
    }
    WB_CATCH
}

int wb_get_chunk_mel(wb_ctx* ctx, int chunk_begin, int n, float* out) {
    WB_TRY
    require_ctx(ctx);
    MelState& s = ctx->mel;
    WB_REQUIRE(s.raw_valid, WB_ESTATE, "no log-mel resident");
    WB_REQUIRE(out && chunk_begin >= 0 && n >= 0 && chunk_begin + n <= s.n_chunks, WB_EINVAL, "chunk range out of bounds");
    const int NM = ctx->cfg.n_mels;
    s.export_buf.reserve((size_t)n * NM * WB_N_FRAMES);
    for (int i = 0; i < n; ++i) {
        const MelChunk& ch = s.h_chunks[chunk_begin + i];
        mel_launch_export(ctx, ch.file, ch.frame_start, WB_N_FRAMES, s.export_buf.p + (size_t)i * NM * WB_N_FRAMES);
    }
    CUDA_CHECK(cudaMemcpyAsync(out, s.export_buf.p, sizeof(float) * (size_t)n * NM * WB_N_FRAMES, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(wb_stream_sync(ctx->stream));
    WB_CATCH
}

int wb_encode(wb_ctx* ctx, const float* mel, int chunk_begin, int B, float* hidden_out) {
    WB_TRY
    require_ctx(ctx);
    const wb_model_cfg& c = ctx->cfg;
    WB_REQUIRE(B >= 1 && B <= c.max_batch, WB_ECAP, "batch %d exceeds max_batch %d", B, c.max_batch);
    const void* in;
    if (mel) {
        const size_t n = (size_t)B * c.n_mels * WB_N_FRAMES;
        CUDA_CHECK(cudaMemcpyAsync(ctx->enc.in_stage.p, mel, sizeof(float) * n, cudaMemcpyHostToDevice, ctx->stream));
        mel_launch_transpose_in(ctx, ctx->enc.in_stage.p, ctx->enc.in_tm.p, B);
        in = ctx->enc.in_tm.p;
    } else {
        WB_REQUIRE(ctx->mel.raw_valid, WB_ESTATE, "wb_encode(mel=NULL) needs a prior wb_log_mel");
        WB_REQUIRE(chunk_begin >= 0 && chunk_begin + B <= ctx->mel.n_chunks, WB_EINVAL, "chunk range out of bounds");
        in = (const char*)ctx->enc.mel_tm.p + (size_t)chunk_begin * (WB_N_FRAMES + 2) * c.n_mels * ctx->esz();
    }
    encoder_forward(ctx, in, B);
    if (hidden_out) {
        CUDA_CHECK(cudaMemcpyAsync(hidden_out, ctx->enc.out.p, sizeof(float) * (size_t)B * c.n_audio_ctx * c.d_model,
                                   cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_CHECK(wb_stream_sync(ctx->stream));
    }
    WB_CATCH
}

int wb_get_encoder_debug(wb_ctx* ctx, const char* what, float* out, int64_t n) {
    WB_TRY
    require_ctx(ctx);
    WB_REQUIRE(ctx->debug, WB_ESTATE, "debug capture is off (wb_set_debug or WB_DEBUG=1)");
    const std::string w(what ? what : "");
    DevBuf<float>* src = w == "stem" ? &ctx->enc.dbg_stem : w == "layer0" ? &ctx->enc.dbg_layer0 : nullptr;
    WB_REQUIRE(src && src->p, WB_EINVAL, "unknown debug tensor '%s'", w.c_str());
    WB_REQUIRE(n <= (int64_t)src->cap, WB_EINVAL, "debug tensor smaller than requested");
    CUDA_CHECK(cudaMemcpyAsync(out, src->p, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(wb_stream_sync(ctx->stream));
    WB_CATCH
}

int wb_greedy_decode(wb_ctx* ctx, int B, const int64_t* prompt, int prompt_len, int max_new_tokens, int64_t eot,
                     const int64_t* suppress, int n_suppress, const int64_t* begin_suppress, int n_begin_suppress,
                     int64_t* tokens_out, int32_t* lens_out, const int64_t* forced, float* logits_out) {
    WB_TRY
    require_ctx(ctx);
    WB_REQUIRE(tokens_out, WB_EINVAL, "null tokens_out");
    DecodeParams p = make_params(B, prompt, prompt_len, max_new_tokens, eot, suppress, n_suppress, begin_suppress,
                                 n_begin_suppress, forced, logits_out != nullptr);
    decoder_run(ctx, p);
    decoder_fetch(ctx, p, tokens_out, lens_out, logits_out);
    WB_CATCH
}

int wb_transcribe_resident(wb_ctx* ctx, const int64_t* prompt, int prompt_len, int max_new_tokens, int64_t eot,
                           const int64_t* suppress, int n_suppress, const int64_t* begin_suppress, int n_begin_suppress,
                           int64_t* tokens_out, int32_t* lens_out, int cap_chunks) {
    WB_TRY
    require_ctx(ctx);
    WB_REQUIRE(tokens_out && lens_out, WB_EINVAL, "null output");
    DecodeParams p = make_params(1, prompt, prompt_len, max_new_tokens, eot, suppress, n_suppress, begin_suppress,
                                 n_begin_suppress, nullptr, false);
    transcribe_resident(ctx, p, tokens_out, lens_out, cap_chunks);
    WB_CATCH
}

int wb_transcribe_batch(wb_ctx* ctx, const float* pcm, const int64_t* offsets, int n_files, const int64_t* prompt,
                        int prompt_len, int max_new_tokens, int64_t eot, const int64_t* suppress, int n_suppress,
                        const int64_t* begin_suppress, int n_begin_suppress, int64_t* tokens_out, int32_t* lens_out,
                        int32_t* file_idx_out, int cap_chunks, int* n_chunks_out) {
    WB_TRY
    require_ctx(ctx);
    WB_REQUIRE(tokens_out && lens_out, WB_EINVAL, "null output");
    int n_chunks = 0;
    upload_pcm(ctx, pcm, offsets, n_files, 0, 0, &n_chunks);
    DecodeParams p = make_params(1, prompt, prompt_len, max_new_tokens, eot, suppress, n_suppress, begin_suppress,
                                 n_begin_suppress, nullptr, false);
    transcribe_resident(ctx, p, tokens_out, lens_out, cap_chunks);
    if (file_idx_out)
        for (int i = 0; i < n_chunks; ++i) file_idx_out[i] = ctx->mel.h_chunks[i].file;
    if (n_chunks_out) *n_chunks_out = n_chunks;
    WB_CATCH
}

}  // extern "C"
