/* whisper_b200.h — C ABI of libwhisper_b200.so: the B200-native replacement for the hot path of
 * KrArunT/whisper-rust-ort's src/main.rs (audio -> log-mel -> encoder -> KV-cache greedy decode
 * -> detokenise).  Plain pointers and sizes only; no C++/torch types cross this boundary.
 *
 * The reference has no FFI of its own: its hot path crosses into third-party
 * code at three `ort::Session::run` sites and is otherwise in-process Rust.  Each entry point
 * below names the reference function / call site it replaces (paths relative to /root/reference).
 * A Rust host binds these with `extern "C"` (INTEGRATION.md shows the stub); here the host above
 * this ABI is C++ (csrc/host/, the `whisper_b200_cli` drop-in) because the image has no Rust.
 *
 * Conventions: every function returning `int` returns WB_OK (0) or a negative WB_E* code and
 * stores a message retrievable with wb_last_error() (thread-local).  Host buffers are caller
 * owned.  A wb_ctx belongs to one GPU and is NOT thread-safe: any host thread may call it, one call at a
 * time (the reference shares `&Session` across rayon threads, main.rs:890-919; here that parallelism is the
 * batch dimension).  Several contexts may be created on one GPU and driven concurrently from different
 * threads (one batch in flight each); contexts of a process with the same device, weight source and
 * architecture share one uploaded copy of the weights.
 * There is no CPU fallback: without a CUDA device every compute call fails with WB_ECUDA.
 */
#ifndef WHISPER_B200_H
#define WHISPER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WB_OK 0
#define WB_EINVAL (-1)   /* bad argument ("Empty audio", main.rs:414-416, maps here) */
#define WB_ECUDA (-2)    /* CUDA runtime error / no device */
#define WB_EIO (-3)      /* file errors */
#define WB_ECAP (-4)     /* exceeds the capacity the ctx was created with */
#define WB_ESTATE (-5)   /* call order error (e.g. decode before encode) */

#define WB_PREC_FP32 0   /* validation build: fp32 weights/activations, SIMT kernels */
#define WB_PREC_BF16 1   /* fast build: bf16 weights + tensor cores, fp32 accumulate */

#define WB_N_FRAMES 3000 /* encoder window, main.rs:896/951 */
#define WB_CHUNK_SAMPLES 480000

typedef struct wb_ctx wb_ctx;

/* Architecture + capacity.  Mirrors whisper-rust-ort_b200/weights.py::ModelCfg (first 9 fields). */
typedef struct wb_model_cfg {
    int32_t n_mels, d_model, n_heads, ffn_dim, enc_layers, dec_layers, vocab, n_audio_ctx, n_text_ctx;
    int32_t precision;     /* WB_PREC_* */
    int32_t max_batch;     /* chunks per encoder/decoder launch (32 = BASELINE config 4) */
    int32_t max_chunks;    /* chunks resident after wb_log_mel (>= max_batch) */
    uint64_t seed;         /* random-init seed when weights_path == NULL */
} wb_model_cfg;

/* CUDA-event timings (ms) of the most recent call of each stage, plus launch counts. */
typedef struct wb_timing {
    float mel_ms, encoder_ms, cross_kv_ms, decode_ms, h2d_ms, d2h_ms;
    int32_t mel_launches, encoder_launches, decode_launches, decode_steps;
} wb_timing;

/* name: "base" | "large-v3" | "toy".  Fills architecture fields + defaults (fp32, batch 32). */
int wb_default_cfg(wb_model_cfg* cfg, const char* name);

/* Replaces ort::init + 3x build_session (main.rs:1090-1108, 169-202): loads weights to HBM.
 * weights_path: a directory holding the reference's ONNX export (encoder_model.onnx +
 * decoder_model.onnx: initializers are read straight from the protobuf, csrc/host/onnx.cpp), a
 * .wb200 blob (weights.py::save_blob), or NULL = seeded random init of the named architecture
 * (BASELINE.json north_star), bit-identical to weights.py::generate(cfg, seed). */
int wb_create(wb_ctx** out, int device, const wb_model_cfg* cfg, const char* weights_path);
void wb_destroy(wb_ctx* ctx);
const char* wb_last_error(void);
/* Number of CUDA devices visible to this process (the CLI's --gpus scheduler maps worker r to device
 * (device + r) % count).  WB_ECUDA when there is none. */
int wb_device_count(int* count_out);
int wb_get_cfg(const wb_ctx* ctx, wb_model_cfg* out);
int wb_get_timing(const wb_ctx* ctx, wb_timing* out);
/* Debug/test: keep copies of encoder intermediates for wb_get_encoder_debug (also WB_DEBUG=1). */
int wb_set_debug(wb_ctx* ctx, int on);
/* How many batches the caller keeps in flight on this GPU (0 = unknown, the default).  The decode GEMMs trade latency
 * against SM-time: with ONE batch in flight (hint 1) they spread over twice the CTAs (one decode chain alone ~6 % faster),
 * with several they keep the fewer, fatter CTAs that cost the other batches less (11 % more throughput at 8 in flight).
 * wb_pool sets the hint of its slots to n_slots.  Results are identical either way. */
int wb_set_load_hint(wb_ctx* ctx, int batches_in_flight);
/* Debug/test: copy a weight tensor (HF state_dict name) back as f32. n = element count. */
int wb_get_tensor(wb_ctx* ctx, const char* name, float* out, int64_t n);

/* ---- group 1: whisper_log_mel_80 (main.rs:407-509) + chunk slicing (:875-882, :895-905) ----
 * pcm: n_files mono 16 kHz files back to back; file i = pcm[offsets[i] .. offsets[i+1]).
 * Per file: reflect-pad 200, periodic Hann, 400-pt FFT hop 160, power, Slaney mel filterbank of cfg.n_mels
 * triangles (80 = the reference's frontend; 128 = the large-v3 one, same function otherwise), log10,
 * clamp to FILE-GLOBAL max-8, (x+4)/4; then the file is cut into 30 s windows every 25 s
 * (chunk_len/step in samples; 0 = 480000/400000) zero-padded IN MEL SPACE to 3000 frames.
 * The chunk batch [n_chunks,n_mels,3000] stays resident on the device for wb_encode.
 * mel_out (nullable, host): per-file [n_mels][floor(N_i/160)] matrices back to back.
 * n_frames_out (nullable, host) [n_files].  Returns chunk count via n_chunks_out (nullable).
 * With mel_out == NULL the call returns once the PCM is on the device and the kernels are enqueued. */
int wb_log_mel(wb_ctx* ctx, const float* pcm, const int64_t* offsets, int n_files,
               int64_t chunk_len, int64_t step, float* mel_out, int64_t* n_frames_out,
               int* n_chunks_out);
/* Same, split for residency benchmarks: upload once, run many times on HBM-resident PCM. */
int wb_upload_pcm(wb_ctx* ctx, const float* pcm, const int64_t* offsets, int n_files,
                  int64_t chunk_len, int64_t step, int* n_chunks_out);
int wb_run_log_mel(wb_ctx* ctx);
/* Resident chunk metadata / data: file index + sample position of each chunk; mel windows. */
int wb_get_chunks(wb_ctx* ctx, int32_t* file_idx, int64_t* sample_pos, int cap);
int wb_get_chunk_mel(wb_ctx* ctx, int chunk_begin, int n, float* out /* [n,n_mels,3000] */);

/* ---- group 2: run_encoder (main.rs:698-707; encoder.run at :703) ----
 * mel: host [B,n_mels,3000] f32, or NULL = use resident chunks [chunk_begin, chunk_begin+B).
 * hidden_out (nullable, host): [B,n_audio_ctx,d_model] f32 (the ONNX output 0).
 * Encoder states and the cross-attention K/V (the `present.*.encoder.*` outputs of
 * decoder_model.onnx, main.rs:786-787) stay resident for wb_greedy_decode.
 * With hidden_out == NULL the call only enqueues (no host wait): wb_greedy_decode runs right behind it on the
 * context's stream, and wb_get_timing waits for whatever is still outstanding. */
int wb_encode(wb_ctx* ctx, const float* mel, int chunk_begin, int B, float* hidden_out);
/* Debug/test: intermediate activations of the last wb_encode. what: "stem" | "layer0". */
int wb_get_encoder_debug(wb_ctx* ctx, const char* what, float* out, int64_t n);

/* ---- group 3: greedy_decode_with_past + argmax_last_dim_raw (main.rs:753-829, 709-735) ----
 * Decodes the B sequences of the last wb_encode.  Loop is held on the device: no host round
 * trip per token.  Semantics kept: step 0 consumes the whole prompt and masks
 * suppress U begin_suppress, later steps mask suppress only; strict '>' argmax (lowest index
 * wins ties, NaN never wins, all-masked -> 0); at most max(1,max_new_tokens) generated ids; a
 * sequence stops after emitting eot (eot is included in its output).
 * tokens_out [B][prompt_len + max(1,max_new_tokens)] (unused tail = -1), lens_out [B].
 * forced (nullable) [B][max(1,max_new_tokens)]: teacher-forced ids fed back instead of the argmax
 * (argmax still reported in tokens_out).  logits_out (nullable) [B][max(1,max_new)][vocab]. */
int wb_greedy_decode(wb_ctx* ctx, int B, const int64_t* prompt, int prompt_len, int max_new_tokens,
                     int64_t eot, const int64_t* suppress, int n_suppress,
                     const int64_t* begin_suppress, int n_begin_suppress,
                     int64_t* tokens_out, int32_t* lens_out,
                     const int64_t* forced, float* logits_out);

/* ---- fused fast path: body of transcribe_longform_chunked (main.rs:870-919) for many files ----
 * log-mel of every file, all chunks batched through encoder + greedy decode in groups of
 * max_batch.  tokens_out [cap_chunks][prompt_len + max(1,max_new)], lens_out/file_idx_out
 * [cap_chunks]; *n_chunks_out = chunks produced (file order, then chunk order). */
int wb_transcribe_batch(wb_ctx* ctx, const float* pcm, const int64_t* offsets, int n_files,
                        const int64_t* prompt, int prompt_len, int max_new_tokens, int64_t eot,
                        const int64_t* suppress, int n_suppress,
                        const int64_t* begin_suppress, int n_begin_suppress,
                        int64_t* tokens_out, int32_t* lens_out, int32_t* file_idx_out,
                        int cap_chunks, int* n_chunks_out);
/* Same on already-uploaded PCM (wb_upload_pcm): the device-resident throughput loop. */
int wb_transcribe_resident(wb_ctx* ctx, const int64_t* prompt, int prompt_len, int max_new_tokens,
                           int64_t eot, const int64_t* suppress, int n_suppress,
                           const int64_t* begin_suppress, int n_begin_suppress,
                           int64_t* tokens_out, int32_t* lens_out, int cap_chunks);

/* ---- batch scheduler behind one handle (north_star (4); replaces the serial per-file loop main.rs:1161-1210) ----
 * A decode is a chain of small dependent kernels, so a GPU reaches its throughput only with several independent
 * batches in flight.  A wb_ctx is ONE batch slot and its calls block; a wb_pool owns n_slots contexts on one device
 * (one shared copy of the weights) plus a worker thread per slot, so that a single-threaded host submits batches
 * without blocking and collects them by ticket.
 * wb_pool_submit: arguments as wb_transcribe_batch; returns a ticket >= 0 (or a negative WB_E* code), blocks only while
 *   n_slots tickets are already waiting for a slot.  pcm / offsets / the output arrays must stay valid until the ticket
 *   has been collected; prompt and suppress lists are copied.
 * wb_pool_wait: blocks until that batch is done, returns its status (the message of a failed batch is then in
 *   wb_last_error() of the calling thread) and the chunks it produced.  Each ticket is collected exactly once, in any order.
 * wb_pool_destroy completes queued work first. */
typedef struct wb_pool wb_pool;
int wb_pool_create(wb_pool** out, int device, const wb_model_cfg* cfg, const char* weights_path, int n_slots);
void wb_pool_destroy(wb_pool* pool);
int wb_pool_slots(const wb_pool* pool);
int wb_pool_submit(wb_pool* pool, const float* pcm, const int64_t* offsets, int n_files,
                   const int64_t* prompt, int prompt_len, int max_new_tokens, int64_t eot,
                   const int64_t* suppress, int n_suppress,
                   const int64_t* begin_suppress, int n_begin_suppress,
                   int64_t* tokens_out, int32_t* lens_out, int32_t* file_idx_out, int cap_chunks);
int wb_pool_wait(wb_pool* pool, int ticket, int* n_chunks_out);

/* ---- measurement hooks (bench.py): CUDA events on the library's own stream ---- */
int wb_mark(wb_ctx* ctx, int slot /* 0..7 */);
int wb_elapsed_ms(wb_ctx* ctx, int slot_a, int slot_b, float* ms_out);   /* syncs on slot_b */
/* Replays one hot kernel `iters` times on the buffers of the last encode/decode (rotating over
 * decoder layers so no launch re-reads what the previous one left in L2) and returns the mean
 * launch duration from CUDA events plus the algorithmic bytes one launch moves.
 * kernel: "cross_attn" | "vocab_proj" | "logmel". */
int wb_bench_kernel(wb_ctx* ctx, const char* kernel, int B, int iters, float* avg_ms_out, double* bytes_per_launch_out);

/* Test hook: runs one bf16 GEMM (A[M,K] row stride lda, W[N,K]) with bias+GELU+f32 residual through the
 * tcgen05 kernel and through the SIMT kernel on the same seeded operands; returns the largest
 * |difference| and the largest |SIMT value| (flags bit0: f32 output instead of bf16). */
int wb_selftest_gemm(wb_ctx* ctx, int M, int N, int K, int lda, int batch, int flags, float* max_diff_out, float* max_abs_out);

/* Test hook (host only, no GPU): the 400-point FFT data flow of the log-mel kernel (20 x 20 four-step, the same
 * mel_math.h code the device runs) on one complex input of 400 points. */
int wb_selftest_fft400(const float* re, const float* im, float* out_re, float* out_im);

/* Test hook (bf16 build): seeded q|k|v for B clips through the tcgen05 attention kernel and the
 * SIMT attention path; returns the largest |difference| and the largest |SIMT value|. */
int wb_selftest_attn(wb_ctx* ctx, int B, float* max_diff_out, float* max_abs_out);

/* ---- host-side pieces of the path (C++ in csrc/host/, exported for the CLI and tests) ---- */
/* load_audio_16k_mono + resample_linear (main.rs:207-316): RIFF/WAVE u8 / s16 / f32 / A-law / mu-law and MPEG-1 / -2 / -2.5
 * Layer III (what symphonia hands the reference as U8, S16 or F32 buffers; s24 / s32 / f64 / ADPCM / FLAC fail with its
 * "Unsupported decoded sample format"), recognised by content like symphonia's probe; channel mean downmix, linear
 * resample to 16 kHz.  *pcm_out is malloc'd; release with wb_host_free. */
int wb_host_load_audio_16k_mono(const char* path, float** pcm_out, int64_t* n_out, double* dur_out);
/* Test hook: an in-memory Layer III stream -> the mono mix the loader builds BEFORE resampling (malloc'd), channels, rate. */
int wb_host_mp3_decode_mono(const uint8_t* bytes, int64_t n, float** out, int64_t* n_out, int* channels, uint32_t* sr);
int64_t wb_host_resample_linear(const float* x, int64_t n, uint32_t sr_in, uint32_t sr_out,
                                float* out, int64_t cap);
void wb_host_free(void* p);
/* Page-locked host memory for PCM staging (H2D copies from it are asynchronous and run at link speed; a pageable
 * source is staged through the driver's bounce buffer first).  `device` = the GPU whose context maps it. */
int wb_host_alloc_pinned(int device, size_t bytes, void** out);
void wb_host_free_pinned(void* p);
/* chunk list of main.rs:875-882; returns count (writes up to cap). */
int wb_host_chunk_starts(int64_t n_samples, int64_t chunk_len, int64_t step, int64_t* out, int cap);
/* stitch_texts / word_overlap (main.rs:659-696). Returns needed length (excluding NUL). */
int wb_host_word_overlap(const char* a, const char* b, int max_words);
int64_t wb_host_stitch_texts(const char* const* chunks, int n, char* out, int64_t cap);
/* Rust's str::to_lowercase as word_overlap applies it to every word (main.rs:687-688): full Unicode lowercase
 * mapping incl. one-to-many images and the Final_Sigma rule.  Returns the needed length (excluding NUL). */
int64_t wb_host_to_lowercase(const char* s, char* out, int64_t cap);
/* percentile / stat_block (main.rs:1021-1048): out6 = {min, median, p90, p95, max, mean}. */
double wb_host_percentile(const double* xs, int n, double p);
int wb_host_stat_block(const double* xs, int n, double* out6);
/* How the output files print an f64 (serde_json = ryu's shortest round-trip digits in its "pretty" layout:
 * "14.884440201999999", "1.0", "0.000482736", "4.82e-6", "1e16", non-finite -> "null"; main.rs:1232-1259).
 * Returns the needed length (excluding NUL), writes up to cap bytes. */
int64_t wb_host_format_f64(double v, char* out, int64_t cap);
/* How the output files print a string: a serde_json string literal (main.rs:1232) and a csv-crate field with
 * QuoteStyle::Necessary (main.rs:1216-1229).  Same return convention. */
int64_t wb_host_json_string(const char* s, char* out, int64_t cap);
int64_t wb_host_csv_field(const char* s, char* out, int64_t cap);

/* tokenizer.json reader + byte-level BPE id->text decoder (replaces the `tokenizers` crate uses:
 * Tokenizer::from_file main.rs:580, token_to_id :531, decode(ids, skip_special=true) :640). */
typedef struct wb_tokenizer wb_tokenizer;
int wb_tokenizer_load(wb_tokenizer** out, const char* tokenizer_json_path);
void wb_tokenizer_free(wb_tokenizer* t);
int64_t wb_tokenizer_token_to_id(const wb_tokenizer* t, const char* token); /* -1 if absent */
/* special_tokens (main.rs:528-569): out5 = {sot, eot, lang, task, no_timestamps}; tok nullable
 * (hard-coded multilingual ids, unknown language/task silently -> en/transcribe). */
int wb_host_special_tokens(const wb_tokenizer* tok, const char* language, const char* task,
                           int64_t* out5);
/* decode_tokens (main.rs:637-648): tok nullable -> "[TOKENS:<first 200 ids>]". Returns needed
 * length (excluding NUL) or negative error. */
int64_t wb_host_decode_tokens(const wb_tokenizer* tok, const int64_t* tokens, int n, char* out,
                              int64_t cap);

/* Host-only test hook for the ONNX-initializer reader: one tensor by Hugging Face parameter name. */
int wb_onnx_read_tensor(const char* onnx_dir, const wb_model_cfg* cfg, const char* name, float* out, int64_t n);

/* The drop-in CLI (main.rs:23-86 flag surface, :1065-1271 driver) as a callable. */
int wb_cli_main(int argc, const char* const* argv);
/* Architecture fields of `cfg` from the HF config.json optimum writes into the export directory (the reference reads shapes
 * from the ONNX graphs themselves; the CLI uses this when --arch is not given). Other fields of `cfg` are left alone. */
int wb_cfg_from_hf_config(const char* config_json_path, wb_model_cfg* cfg);

#ifdef __cplusplus
}
#endif
#endif /* WHISPER_B200_H */
