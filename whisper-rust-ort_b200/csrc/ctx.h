// ctx.h — the per-GPU context behind the opaque `wb_ctx` of include/whisper_b200.h, and the
// internal interfaces between the translation units (mel / gemm / encoder / decoder / weights).
#pragma once
#include <cmath>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "common.h"

// ---------------- log-mel ----------------
struct MelTables {                 // POD, copied to the device once
    float window[400];             // periodic Hann, main.rs:323-330
    float tw_re[400], tw_im[400];  // W_400^k
    float fb_w[512];               // non-zero filterbank weights, mel-major runs
    int fb_idx[384];               // start[n_mels], len[n_mels], offset-into-fb_w[n_mels]  (n_mels = 80 or 128)
};
struct MelTile {                   // one 24-frame tile of K1a (built on the device by mel_tiles_kernel)
    int64_t pcm_off, N, raw_base;  // file start in the PCM buffer, file length, first log-mel row of the tile
    int f0, nf, file, pad;         // first frame of the tile within the file, frames of the file, file index
};
struct MelChunk {
    int file;
    int frame_start;               // chunk_pos / 160 (main.rs:895)
};
struct MelState {
    DevBuf<float> pcm;
    DevBuf<int64_t> file_off, frame_off;
    DevBuf<int> tile_off, fmax;
    DevBuf<MelTile> tiles;
    int n_files_staged = 0;        // files whose offsets are on the device (= n_files once the upload is committed)
    DevBuf<float> raw;             // [total_frames][n_mels] log10 mel, time-major
    DevBuf<MelChunk> chunks;
    DevBuf<float> export_buf;
    std::vector<int64_t> h_file_off, h_frame_off, h_chunk_pos;
    std::vector<int> h_tile_off;
    std::vector<MelChunk> h_chunks;
    int n_files = 0, n_chunks = 0, total_tiles = 0;
    int64_t total_frames = 0;
    bool raw_valid = false;
};

// ---------------- weights ----------------
struct LinearW {
    void* w = nullptr;             // [out][in] in the compute dtype (f32 or bf16)
    float* b = nullptr;            // [out] f32 or nullptr
    int out = 0, in = 0;
};
struct LNW {
    float* w = nullptr;
    float* b = nullptr;
};
struct EncLayerW {
    LNW ln1, ln2;
    LinearW qkv, o, fc1, fc2;
};
struct DecLayerW {
    LNW ln1, ln2, ln3;
    LinearW qkv, o, cq, ckv, co, fc1, fc2;
};
struct ModelW {
    LinearW conv1, conv2;          // packed [d][3*C_in], k-major then channel
    float* enc_pos = nullptr;      // [n_audio_ctx][d] f32
    std::vector<EncLayerW> enc;
    LNW enc_ln;
    void* embed = nullptr;         // [vocab][d] compute dtype (tied input embedding / output proj)
    float* dec_pos = nullptr;      // [n_text_ctx][d] f32
    std::vector<DecLayerW> dec;
    LNW dec_ln;
    std::shared_ptr<struct WeightStore> store;        // owns the device allocations + f32 originals; shared by every
                                                      // context of the process with the same device/source/config
};
// One uploaded copy of a model per (device, source, architecture, precision): contexts in flight on a GPU read
// the same weights (read-only during runs), so the second and later wb_create skip generation/parse + upload.
struct WeightStore {
    int device = 0;
    std::vector<void*> allocs;
    std::map<std::string, std::vector<float>> host;   // f32 originals by HF name
    ModelW view;                                      // the pointer table (view.store stays empty)
    ~WeightStore();
};

// ---------------- activations ----------------
struct EncBufs {
    DevBuf<unsigned char> mel_tm;      // [max_chunks][3002][n_mels] compute dtype (resident chunks)
    DevBuf<unsigned char> in_tm;       // [max_batch][3002][n_mels]   (host-provided mel path)
    DevBuf<float> in_stage;            // [max_batch][n_mels][3000] f32 staging of host mel
    DevBuf<unsigned char> h1p;         // [B][3001][d] compute dtype, row 0 zero
    DevBuf<float> x;                   // [B][1500][d] residual stream
    DevBuf<unsigned char> h;           // [B][1500][d] LN output
    DevBuf<unsigned char> qkv;         // [B][1500][3d]
    DevBuf<unsigned char> att;         // [B][1500][d]
    DevBuf<unsigned char> ffn;         // [B][1500][ffn]
    DevBuf<float> scores;              // [G*H][1500][1500] (SIMT attention path)
    DevBuf<float> out;                 // [B][1500][d] f32 encoder output (ONNX output 0)
    DevBuf<unsigned char> out_c;       // compute-dtype copy (bf16 mode)
    DevBuf<unsigned char> ckv;         // [dec_layers][B][1500][2d] compute dtype
    DevBuf<float> dbg_stem, dbg_layer0;
    int attn_group = 4;
    int B_valid = 0;                   // sequences encoded by the last wb_encode
};
struct DecGraph {
    int key[8];
    cudaGraphExec_t exec;
    int launches;
};
struct DecBufs {
    DevBuf<float> x, qkv, att, q, ffn, logits;
    DevBuf<unsigned char> self_kv;     // [dec_layers][B][T_max][2d] compute dtype
    DevBuf<float> logits_all;          // optional [B][steps][V]
    DevBuf<int> tokens;                // [B][T_total] generated+prompt ids (device)
    DevBuf<int> forced;                // [B][max_new] teacher forcing (optional)
    DevBuf<int> lens, finished, state; // state: [0]=position t, [1]=#unfinished
    DevBuf<float> amax_buf;            // fused arg-max partials of the vocabulary projection
    float* amax_val = nullptr; int* amax_idx = nullptr; int* amax_state = nullptr; int amax_ctas = 0;
    bool fuse_argmax = true, want_logits = false;
    DevBuf<unsigned int> sup_base, sup_first;   // vocab bitmaps
    int T_max = 0;
    int last_T_total = 0;              // prompt_len + max_new of the last decode (layout of `tokens`)
    // decode-segment CUDA graphs (prompt prefix, SEG generated tokens, remainder), cached by key
    std::vector<DecGraph> graphs;
    int* unfinished_host = nullptr;    // mapped: sequences still running, written at the end of a segment
    int* unfinished_dev = nullptr;
    bool pdl = false;                  // programmatic dependent launch for the decode chain
    bool ffn_handoff = true;           // fc1 writes its output as bf16 for fc2 (bit-identical, half the staging bytes)
    int lean = 0;                      // 1: register-capped skinny GEMMs + 4-warp cross-attention (see decoder.cu)
    int load_hint = 0;                 // wb_set_load_hint: batches the caller keeps in flight (0 unknown); 1 selects the latency-oriented GEMM shapes
    int self_attn_warps = 4;           // warps per (sequence, head) in the self-attention kernel (WB_SELF_ATTN_WARPS = 2 | 4 | 8)
    int* stage_host = nullptr;         // pinned staging of the per-decode control state (ids, bitmaps, lens): a pageable
    size_t stage_ints = 0;             //   source would make every cudaMemcpyAsync wait for the stream to drain first
    void* vocab_tc = nullptr;          // vocab_tc.cu: tcgen05 vocabulary projection + arg-max (bf16 build, d_model <= 512)
    struct DecCluster* cluster = nullptr;   // dec_cluster.cu: all layers of a step in one launch (bf16, whisper-base widths)
};

struct wb_ctx {
    wb_model_cfg cfg{};
    int device = 0;
    cudaStream_t stream = nullptr;
    ModelW w;
    MelTables mel_tables{};
    MelTables* mel_tables_dev = nullptr;
    MelState mel;
    EncBufs enc;
    DecBufs dec;
    wb_timing timing{};
    bool debug = false;
    int sm_count = 148;
    CudaEvent ev0, ev1;
    // stage timings are read back lazily (timing_flush): no host wait between log-mel, encoder and decode
    CudaEvent mel_e0, mel_e1, enc_e0, enc_e1, enc_e2;
    bool mel_t_pending = false, enc_t_pending = false;
    CudaEvent marks[8];
    size_t esz() const { return cfg.precision == WB_PREC_BF16 ? 2 : 4; }
};

// vocab_tc.cu — final LayerNorm + vocabulary projection on tcgen05 with the masked arg-max fused (swap-AB: 128 weight rows x
// <= 32 sequences per MMA); returns the number of per-CTA partials written for argmax_merge_kernel
void vocab_tc_alloc(wb_ctx* ctx);
void vocab_tc_free(wb_ctx* ctx);
bool vocab_tc_ok(const wb_ctx* ctx, int B);
int vocab_tc_stride(int B);          // sequences per partial row ([cta][stride] values | indices)
int vocab_tc_launch(wb_ctx* ctx, cudaStream_t st, bool pdl, const int* state, const float* x, int B, float* amax_val, int* amax_idx);
// the per-layer decode GEMMs on the same kernel family; false = shape / weight not served, caller falls back
bool skinny_tc_launch(wb_ctx* ctx, cudaStream_t st, bool pdl, const float* X, int B, int K, const void* W, int N, const float* bias,
                      const float* ln_w, const float* ln_b, int act, const float* residual, float* Y);

// api.cpp — waits for and publishes the stage timings whose events are still outstanding
void timing_flush(wb_ctx* ctx);

// mel.cu
#define WB_MEL_FPT 24                 // frames per log-mel tile (mel.cu); wb_upload_pcm builds the tile prefix sums with it
void mel_build_tables(MelTables& t, int n_mels);
void mel_set_attrs();
int64_t mel_n_frames(int64_t n);
void mel_build_tiles(wb_ctx* ctx);
void mel_launch_raw(wb_ctx* ctx);
void mel_launch_chunks(wb_ctx* ctx, int chunk0, int n, void* out);
void mel_launch_export(wb_ctx* ctx, int file, int64_t frame0, int64_t n_out, float* out_dev);
void mel_launch_transpose_in(wb_ctx* ctx, const float* in_dev, void* out, int B);

// weights.cpp
void weights_init(wb_ctx* ctx, const char* path);
void weights_free(wb_ctx* ctx);

// gemm_simt.cu — C = epilogue(alpha * A[M,K] * B^T) ; B is [N][K] (K contiguous) or, if b_kn,
// [K][N] (N contiguous).  Batched over z with a two-level (outer, inner) stride decomposition.
enum { WB_F32 = 0, WB_BF16 = 1 };
struct GemmArgs {
    const void* A = nullptr; const void* B = nullptr; void* C = nullptr;
    int ta = WB_F32, tb = WB_F32, tc = WB_F32;
    int M = 0, N = 0, K = 0, lda = 0, ldb = 0, ldc = 0;
    int batch = 1, inner = 1;
    long long sAo = 0, sAi = 0, sBo = 0, sBi = 0, sCo = 0, sCi = 0;
    bool b_kn = false;
    float alpha = 1.0f;
    const float* bias = nullptr;        // [N]
    int act = 0;                        // 1 = exact-erf GELU
    const float* rowadd = nullptr;      // [M][N] f32 added after activation (positional table)
    int ld_rowadd = 0;
    const float* residual = nullptr;    // f32, same indexing as C (may alias C)
};
void gemm_simt(wb_ctx* ctx, const GemmArgs& a);
// gemm_tc.cu — tcgen05/TMEM/TMA kernel (bf16 operands, K-major, N % 128 == 0)
void gemm_tc_set_attrs();
bool gemm_tc_eligible(const GemmArgs& a);
void gemm_tc(wb_ctx* ctx, const GemmArgs& a);
// dispatcher: tensor-core kernel when eligible, else SIMT
void gemm(wb_ctx* ctx, const GemmArgs& a);

// attn_tc.cu — tcgen05 flash attention (bf16 build)
void attn_tc_set_attrs();
bool attn_tc_enabled();
void attn_tc(wb_ctx* ctx, const void* qkv, void* out, int B, int T, int d, int H);
// encoder.cu
int attention_simt(wb_ctx* ctx, const void* qkv, void* att, int B);
void encoder_alloc(wb_ctx* ctx);
void encoder_forward(wb_ctx* ctx, const void* mel_tm, int B);   // mel_tm: [B][3002][n_mels] compute dtype
// dec_cluster.cu — embedding + all decoder layers of one step as ONE cluster-chained launch
void dec_cluster_alloc(wb_ctx* ctx);
void dec_cluster_free(wb_ctx* ctx);
bool dec_cluster_enabled(const wb_ctx* ctx);
bool dec_cluster_vocab_ok(const wb_ctx* ctx, int B);
// stand-alone cross-attention on the TMA ring + tensor cores (bf16 build, head_dim 64)
bool cross_attn_tc_ok(const wb_ctx* ctx);
void cross_attn_tc(wb_ctx* ctx, cudaStream_t st, bool pdl, int layer, const float* q, float* out, int B);
void dec_cluster_vocab(wb_ctx* ctx, cudaStream_t st, bool pdl, int* state, int B, float* logits, const int* forced, int max_new, int eot,
                       int T_total, int* cur_tok);
double dec_cluster_vocab_bytes(const wb_ctx* ctx);
double dec_cluster_bytes(const wb_ctx* ctx, int B);
void dec_cluster_print_prof(wb_ctx* ctx);
void dec_cluster_layers(wb_ctx* ctx, cudaStream_t st, bool pdl, const int* state, const int* prompt_dev, const int* cur_tok, int B);
// decoder.cu
void decoder_alloc(wb_ctx* ctx);
struct DecodeParams {
    int B, prompt_len, max_new;
    int eot;
    const int64_t* prompt;
    const int64_t* suppress; int n_suppress;
    const int64_t* begin_suppress; int n_begin_suppress;
    const int64_t* forced;
    bool want_logits;
};
void decoder_run(wb_ctx* ctx, const DecodeParams& p);           // async on ctx->stream
void decoder_bench(wb_ctx* ctx, const char* kernel, int B, int iters, float* avg_ms, double* bytes);
void decoder_fetch(wb_ctx* ctx, const DecodeParams& p, int64_t* tokens_out, int32_t* lens_out, float* logits_out);
