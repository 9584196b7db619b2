/* oracle/mel_ref.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, f32 arithmetic) of the reference's log-mel frontend,
 * /root/reference/src/main.rs:323-509, used only by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py as the checker for the CUDA kernels in
 * whisper-rust-ort_b200/csrc/mel.cu.  Nothing under whisper-rust-ort_b200/ links or calls this.
 *
 * Parity pinning: the reference holds NO numeric golden vectors for this path (SURVEY.md §8c,
 * "parity unpinned" by the reference itself).  This restatement is pinned instead against
 *   (1) tests/golden/mel_hf_*.npz — outputs of transformers.WhisperFeatureExtractor (the
 *       implementation the reference's comments say it approximates, main.rs:318-322, 418, 449)
 *       on exact-30 s clips, where both semantics coincide, and
 *   (2) its own f64 direct-DFT mode (use_dft64=1), which removes FFT rounding from the picture.
 *
 * The one third-party piece on this path is the 400-point FFT: crate rustfft 6.4.1
 * (Cargo.lock:991; call sites main.rs:440-441, 473), absent from /root/reference.  Its published
 * algorithm for a composite length is an f32 mixed-radix Cooley-Tukey with twiddles evaluated in
 * f64 and rounded to f32; fft400_f32() below restates that (radices 4,4,5,5).  Results differ
 * from rustfft at the f32-rounding level (~1e-7 relative), far inside the 1e-4 tolerance.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define N_FFT 400
#define HOP 160
#define N_MELS 80          /* the reference's frontend; the _n entry points take 80 or 128 (large-v3) */
#define N_MELS_MAX 128
#define N_FREQ 201
#define SR 16000

typedef struct { float re, im; } cpx;

/* main.rs:323-330 — periodic Hann, 2*pi*i formed in f32 then divided by n */
void wbref_hann(float* w, int n) {
    for (int i = 0; i < n; ++i) {
        float x = (3.14159265358979323846f * 2.0f * (float)i) / (float)n;
        w[i] = 0.5f - 0.5f * cosf(x);
    }
}

/* main.rs:332-341 */
static float hz_to_mel_slaney(float hz) {
    const float min_log_hz = 1000.0f, min_log_mel = 15.0f;
    const float logstep = 27.0f / logf(6.4f);
    float mel = 3.0f * hz / 200.0f;
    if (hz >= min_log_hz) mel = min_log_mel + logf(hz / min_log_hz) * logstep;
    return mel;
}

/* main.rs:343-352 */
static float mel_to_hz_slaney(float mel) {
    const float min_log_hz = 1000.0f, min_log_mel = 15.0f;
    const float logstep = logf(6.4f) / 27.0f;
    float hz = 200.0f * mel / 3.0f;
    if (mel >= min_log_mel) hz = min_log_hz * expf(logstep * (mel - min_log_mel));
    return hz;
}

/* main.rs:354-405 — fb is [n_mels][n_freq] row-major */
void wbref_mel_filterbank_n(float* fb, int n_mels) {
    const int n_freq = N_FREQ;
    float fmax = fminf(8000.0f, (float)SR / 2.0f);
    float mel_min = hz_to_mel_slaney(0.0f), mel_max = hz_to_mel_slaney(fmax);
    float freq_points[N_MELS_MAX + 2], fft_freqs[N_FREQ];
    for (int i = 0; i < n_mels + 2; ++i) {
        float m = mel_min + (mel_max - mel_min) * (float)i / (float)(n_mels + 1);
        freq_points[i] = mel_to_hz_slaney(m);
    }
    float max_hz = (float)SR / 2.0f;
    for (int k = 0; k < n_freq; ++k) fft_freqs[k] = (float)k * max_hz / (float)(n_freq - 1);
    for (int m = 0; m < n_mels; ++m) {
        float f_left = freq_points[m], f_center = freq_points[m + 1], f_right = freq_points[m + 2];
        float denom_left = fmaxf(f_center - f_left, 1e-6f);
        float denom_right = fmaxf(f_right - f_center, 1e-6f);
        for (int k = 0; k < n_freq; ++k) {
            float f = fft_freqs[k];
            float lower = (f - f_left) / denom_left;
            float upper = (f_right - f) / denom_right;
            fb[m * n_freq + k] = fmaxf(fminf(lower, upper), 0.0f);
        }
    }
    for (int m = 0; m < n_mels; ++m) {
        float enorm = 2.0f / fmaxf(freq_points[m + 2] - freq_points[m], 1e-6f);
        for (int k = 0; k < n_freq; ++k) fb[m * n_freq + k] *= enorm;
    }
}

void wbref_mel_filterbank(float* fb) { wbref_mel_filterbank_n(fb, N_MELS); }

/* ---- 400-point forward FFT, f32 mixed radix (restating rustfft's approach) ---- */
static cpx g_tw[N_FFT];
static double g_twd[N_FFT][2];
static int g_tw_ready = 0;
static void init_tw(void) {
    if (g_tw_ready) return;
    for (int k = 0; k < N_FFT; ++k) {
        double a = -2.0 * 3.14159265358979323846 * (double)k / (double)N_FFT;
        g_twd[k][0] = cos(a); g_twd[k][1] = sin(a);
        g_tw[k].re = (float)g_twd[k][0]; g_tw[k].im = (float)g_twd[k][1];
    }
    g_tw_ready = 1;
}
static inline cpx cmul(cpx a, cpx b) {
    cpx r = { a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re };
    return r;
}
static void fft_rec(const cpx* in, cpx* out, int n, int stride) {
    if (n == 1) { out[0] = in[0]; return; }
    int r = (n % 4 == 0) ? 4 : (n % 5 == 0) ? 5 : (n % 2 == 0) ? 2 : n;
    int m = n / r;
    for (int j = 0; j < r; ++j) fft_rec(in + (size_t)j * stride, out + (size_t)j * m, m, stride * r);
    cpx t[8], y[8];
    for (int k = 0; k < m; ++k) {
        for (int j = 0; j < r; ++j) t[j] = cmul(out[j * m + k], g_tw[(j * k * (N_FFT / n)) % N_FFT]);
        for (int p = 0; p < r; ++p) {
            cpx acc = t[0];
            for (int j = 1; j < r; ++j) {
                cpx w = g_tw[((j * p) % r) * (N_FFT / r)];
                cpx v = cmul(t[j], w);
                acc.re += v.re; acc.im += v.im;
            }
            y[p] = acc;
        }
        for (int p = 0; p < r; ++p) out[k + m * p] = y[p];
    }
}
void wbref_fft400_f32(const float* re_in, const float* im_in, float* re_out, float* im_out) {
    init_tw();
    cpx a[N_FFT], b[N_FFT];
    for (int i = 0; i < N_FFT; ++i) { a[i].re = re_in[i]; a[i].im = im_in ? im_in[i] : 0.0f; }
    fft_rec(a, b, N_FFT, 1);
    for (int i = 0; i < N_FFT; ++i) { re_out[i] = b[i].re; im_out[i] = b[i].im; }
}

long wbref_n_frames(long n) {
    /* main.rs:444-452: 1 + (padded - win)/hop, minus the dropped last frame */
    long padded = n + N_FFT;
    long nf = padded < N_FFT ? 1 : 1 + (padded - N_FFT) / HOP;
    if (nf > 1) nf -= 1;
    return nf;
}

/* main.rs:407-509.  out is [80][n_frames] row-major. Returns 0, or -1 on "Empty audio".
 * use_dft64 != 0 replaces the f32 FFT by an f64 direct DFT (validation mode).
 * If raw_log10 != NULL it also receives log10(max(mel,1e-10)) before clamp/scale, and
 * *gmax_out the file-global maximum (what the two-phase CUDA kernel exchanges). */
int wbref_log_mel_n(const float* audio, long n, float* out, int n_mels, int use_dft64, float* raw_log10,
                    float* gmax_out) {
    if (n <= 0) return -1;
    if (n_mels < 1 || n_mels > N_MELS_MAX) return -2;
    init_tw();
    const int pad = N_FFT / 2;
    long plen = n + 2 * pad;
    float* padded = (float*)calloc((size_t)plen, sizeof(float));
    if (n >= 2) {
        for (int i = 0; i < pad; ++i) {
            long idx = pad - i;
            long src = idx < n - 1 ? idx : n - 1;
            padded[i] = audio[src];
        }
        memcpy(padded + pad, audio, (size_t)n * sizeof(float));
        for (int i = 0; i < pad; ++i) {
            long idx = n - 2 - i; if (idx < 0) idx = 0;          /* saturating_sub */
            padded[pad + n + i] = audio[idx];
        }
    } else {
        memcpy(padded, audio, (size_t)n * sizeof(float));         /* then zero-resized */
    }
    float window[N_FFT];
    wbref_hann(window, N_FFT);
    float* fb = (float*)malloc(sizeof(float) * n_mels * N_FREQ);
    wbref_mel_filterbank_n(fb, n_mels);
    long n_frames = wbref_n_frames(n);

    cpx fin[N_FFT], fout[N_FFT];
    float pows[N_FREQ];
    for (long frame = 0; frame < n_frames; ++frame) {
        long start = frame * HOP;
        for (int i = 0; i < N_FFT; ++i) {
            long idx = start + i;
            float s = idx < plen ? padded[idx] : 0.0f;
            fin[i].re = s * window[i]; fin[i].im = 0.0f;
        }
        if (!use_dft64) {
            fft_rec(fin, fout, N_FFT, 1);
            for (int k = 0; k < N_FREQ; ++k) pows[k] = fout[k].re * fout[k].re + fout[k].im * fout[k].im;
        } else {
            for (int k = 0; k < N_FREQ; ++k) {
                double re = 0.0, im = 0.0;
                for (int i = 0; i < N_FFT; ++i) {
                    int t = (int)(((long)k * i) % N_FFT);
                    re += (double)fin[i].re * g_twd[t][0];
                    im += (double)fin[i].re * g_twd[t][1];
                }
                float fr = (float)re, fi = (float)im;
                pows[k] = fr * fr + fi * fi;
            }
        }
        for (int m = 0; m < n_mels; ++m) {
            float e = 0.0f;
            for (int k = 0; k < N_FREQ; ++k) e += fb[m * N_FREQ + k] * pows[k];
            out[(size_t)m * n_frames + frame] = fmaxf(e, 1e-10f);
        }
    }
    float max_log = -INFINITY;
    size_t total = (size_t)n_mels * (size_t)n_frames;
    for (size_t i = 0; i < total; ++i) {
        float lv = log10f(out[i]);
        if (lv > max_log) max_log = lv;
    }
    if (gmax_out) *gmax_out = max_log;
    for (size_t i = 0; i < total; ++i) {
        float lv = log10f(out[i]);
        if (raw_log10) raw_log10[i] = lv;
        float clamped = fmaxf(lv, max_log - 8.0f);
        out[i] = (clamped + 4.0f) / 4.0f;
    }
    free(fb); free(padded);
    return 0;
}
/* whisper_log_mel_80 itself */
int wbref_log_mel(const float* audio, long n, float* out, int use_dft64, float* raw_log10, float* gmax_out) {
    return wbref_log_mel_n(audio, n, out, N_MELS, use_dft64, raw_log10, gmax_out);
}

/* Batch of equal-length clips, pthreads over clips (cpu_baseline with all host threads;
 * this image has no libgomp). */
#include <pthread.h>
typedef struct { const float* audio; long n_clips, n, nf; float* out; long* next; pthread_mutex_t* mu; int rc; } job_t;
static void* worker(void* p) {
    job_t* j = (job_t*)p;
    for (;;) {
        pthread_mutex_lock(j->mu);
        long c = (*j->next)++;
        pthread_mutex_unlock(j->mu);
        if (c >= j->n_clips) break;
        int r = wbref_log_mel(j->audio + c * j->n, j->n, j->out + (size_t)c * N_MELS * j->nf, 0, NULL, NULL);
        if (r) j->rc = r;
    }
    return NULL;
}
int wbref_log_mel_batch(const float* audio, long n_clips, long n, float* out, int n_threads) {
    long nf = wbref_n_frames(n), next = 0;
    init_tw();
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
    pthread_t th[256]; job_t jobs[256];
    for (int t = 0; t < n_threads; ++t) {
        job_t j = { audio, n_clips, n, nf, out, &next, &mu, 0 };
        jobs[t] = j;
        pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    int rc = 0;
    for (int t = 0; t < n_threads; ++t) { pthread_join(th[t], NULL); if (jobs[t].rc) rc = jobs[t].rc; }
    return rc;
}
