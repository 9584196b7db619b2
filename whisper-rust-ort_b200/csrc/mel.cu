// mel.cu — kernel group 1: batched Whisper log-mel (replaces whisper_log_mel_80,
// /root/reference/src/main.rs:407-509, and the chunk slicing of :895-905).
//
// K1a logmel_raw_kernel : PCM -> log10(max(mel,1e-10)) time-major [frame][n_mels] + per-FILE max
//                         (the reference clamps against the max of the whole file, quirk Q1).
// K1b mel_chunks_kernel : clamp(max-8), (x+4)/4, cut 3000-frame windows, zero-pad in mel space
//                         (Q2) -> [chunk][1+3000+1][n_mels] time-major in the encoder's compute
//                         dtype, the layout the conv-stem GEMM consumes without im2col.
// K1c mel_export_kernel : same normalisation, reference layout [n_mels][frames] f32 (API/test only).
// n_mels is 80 (the reference's whisper_log_mel_80) or 128 (the large-v3 frontend: same code, filterbank of 128
// triangles over the same 201 bins — BASELINE.json configs[4]).
//
// HBM-bound by design (1.92 MB in + 0.96 MB out per 30 s clip); the FFT is a
// register-resident 20x20 four-step (mel_math.h) so the only global traffic is PCM in, mel out.
#include "ctx.h"
#include "mel_math.h"
#define WB_PACKED_F32X2
#include "mel_math.h"
#undef WB_PACKED_F32X2

namespace {

constexpr int FPT = WB_MEL_FPT;               // frames per tile (24: 3000 frames = 125 tiles exactly)
constexpr int NPAIR = FPT / 2;                // two real frames ride one complex FFT
constexpr int MEL_THREADS = NPAIR * 20;       // thread = (frame pair, FFT column): every lane has a column (15 half-warps)
constexpr int SPAN = (FPT - 1) * 160 + 400;   // padded samples a tile touches
constexpr int SKEW = 20;                      // sample j sits at j + SKEW*(j/320): pair p's column reads hit bank tid%32
constexpr int PCM_LD = SPAN + SKEW * ((SPAN + 319) / 320);
constexpr int SCR = 420;                      // complex per FFT: 20 x 21 (padded) rows; 420 = 4 mod 16 keeps (pair, column) lanes apart
constexpr int PPL = 230;                      // power-spectrum row pitch = 6 mod 32: the six frame groups of a warp sit 6 banks apart
constexpr int FBW_MAX = 512;                  // non-zero filterbank weights (<= 2 per FFT bin + slack)
static_assert(FPT * PPL + FPT * (128 + 6) <= 2 * NPAIR * SCR, "power rows and the output tile live in the FFT scratch");
// frame f of a tile = 4*fg + q (frame group fg, lane-parallel; q = 0..3, register-parallel) lives in row fg + 6*q of the
// power and output tiles: the lanes of a warp then walk consecutive rows, 6 banks apart, each up to 6 mel bins wide
__device__ __forceinline__ int tile_row(int f) { return (f >> 2) + 6 * (f & 3); }

__device__ __forceinline__ float padded_sample(const float* __restrict__ x, int64_t N, int64_t p) {
    // main.rs:419-435: reflect-pad 200 each side (N >= 2), else audio then zeros.
    if (N >= 2) {
        int64_t n = p - 200;
        if (n < 0) {
            int64_t idx = 200 - p;
            return x[idx < N - 1 ? idx : N - 1];
        }
        if (n < N) return x[n];
        int64_t i = n - N;
        if (i >= 200) return 0.0f;                       // main.rs:468 (never hit for valid frames)
        int64_t idx = N - 2 - i;
        return x[idx > 0 ? idx : 0];
    }
    return p < N ? x[p] : 0.0f;
}

__device__ __forceinline__ void atomic_max_float(int* addr, float v) {
    if (v >= 0.0f) atomicMax(addr, __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

__device__ __forceinline__ void cp_async_4(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" :: "r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc));
}

// tile -> (file, first frame, where its log-mel rows go): built once per upload, so K1a does one load per tile
// instead of a binary search (a chain of dependent global loads) in front of every tile.
__global__ void mel_tiles_kernel(const int64_t* __restrict__ file_off, const int64_t* __restrict__ frame_off,
                                 const int* __restrict__ tile_off, int n_files, int total_tiles, MelTile* __restrict__ out) {
    const int tile = blockIdx.x * blockDim.x + threadIdx.x;
    if (tile >= total_tiles) return;
    int lo = 0, hi = n_files - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (tile_off[mid] <= tile) lo = mid; else hi = mid - 1;
    }
    MelTile t;
    t.pcm_off = file_off[lo];
    t.N = file_off[lo + 1] - file_off[lo];
    t.f0 = (tile - tile_off[lo]) * FPT;
    t.nf = (int)(frame_off[lo + 1] - frame_off[lo]);
    t.raw_base = frame_off[lo] + t.f0;
    t.file = lo;
    t.pad = 0;
    out[tile] = t;
}

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 100000;\n"
        "selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done != 0;
}

// Stage the padded PCM span of a tile (skewed layout: 320-sample regions, 340 floats apart) while the previous
// tile's FFTs run.  Interior 16-byte-aligned tiles (all but the first / last of a file) are 13 bulk copies issued by
// one thread and written by the TMA unit (no load/store-unit wavefronts; completion on `bar`); interior tiles of a
// file that starts off a 16-byte boundary use 4-byte cp.async; file edges take the reflecting scalar path.
// Returns true when the data arrives on the mbarrier (uniform over the CTA).
__device__ __forceinline__ bool stage_pcm(float* s_pcm, uint32_t bar, const float* __restrict__ pcm, const MelTile& t, int tid) {
    const int64_t p0 = (int64_t)t.f0 * 160;                      // first padded sample of the tile
    const float* x = pcm + t.pcm_off;
    const float* src = x + (p0 - 200);
    if (p0 >= 200 && p0 - 200 + SPAN <= t.N) {
        if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
            if (tid == 0) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(SPAN * 4)) : "memory");
#pragma unroll
                for (int r = 0; r < (SPAN + 319) / 320; ++r) {
                    const uint32_t bytes = (uint32_t)((SPAN - 320 * r < 320 ? SPAN - 320 * r : 320) * 4);
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(smem_addr(s_pcm + r * (320 + SKEW))), "l"(src + 320 * r), "r"(bytes), "r"(bar) : "memory");
                }
            }
            return true;
        }
        for (int j = tid; j < SPAN; j += MEL_THREADS) cp_async_4(s_pcm + j + SKEW * (j / 320), src + j);
    } else {
        for (int j = tid; j < SPAN; j += MEL_THREADS) s_pcm[j + SKEW * (j / 320)] = padded_sample(x, t.N, p0 + j);
    }
    asm volatile("cp.async.commit_group;\n" ::);
    return false;
}

// K1a.  Persistent CTAs (three per SM) walk the tiles of 24 frames.  Per tile: thread (pair, column) runs the 20-point
// column DFT of the pair's packed complex frame out of the staged PCM, twiddles and scatters it; the next tile's PCM
// starts streaming in; the same thread runs the 20-point row DFT; the Hermitian split gives both power spectra; the
// sparse mel filterbank runs with a warp's lanes on the same few mel bins of different frame groups (weights
// broadcast, no divergence on the run length); the [24][n_mels] log10 tile goes out in coalesced stores.
// Shared-memory bandwidth and latency bound this kernel, so every access pattern is laid out conflict-free (SKEW, SCR,
// the [k1][column] twiddle table, the rotated spectrum base, PPL, the odd output pitch).
template <int NM, bool PACKED>
__global__ void __launch_bounds__(MEL_THREADS, 3)
logmel_raw_kernel(const float* __restrict__ pcm, const MelTile* __restrict__ tiles, int total_tiles,
                  const MelTables* __restrict__ tab, float* __restrict__ raw, int* __restrict__ fmax) {
    constexpr int OP = NM + (6 - NM % 32 + 32) % 32;           // output tile pitch = 6 mod 32 (80 -> 102, 128 -> 134)
    static_assert(OP % 32 == 6 && FPT * PPL + FPT * OP <= 2 * NPAIR * SCR, "output tile");
    extern __shared__ __align__(16) float smem[];
    float* s_pcm = smem;                                       // PCM_LD (skewed)
    float* s_win = s_pcm + PCM_LD;                             // 400
    float* s_twr = s_win + 400;                                // 400 + 400: W_400^(k1*col) at [k1*20 + col], re and im planes (two
    float* s_twi = s_twr + 400;                                //   32-bit loads: lanes of different pairs broadcast, columns never collide)
    float* s_fbw = s_twi + 400;                                // FBW_MAX, pre-scaled by 1/4 (the Hermitian split's 1/2, squared)
    int* s_fbi = reinterpret_cast<int*>(s_fbw + FBW_MAX);      // start[NM], len[NM], off[NM]
    float2* s_scr = reinterpret_cast<float2*>(s_fbi + 3 * NM + ((3 * NM) & 1));   // NPAIR * SCR complex
    float* s_pow = reinterpret_cast<float*>(s_scr);            // later: FPT power rows of PPL ...
    float* s_out = s_pow + FPT * PPL;                          // ... and the [FPT][OP] output tile behind them
    __shared__ int s_tmax;
    __shared__ __align__(8) uint64_t s_bar;                    // PCM of a bulk-copied tile has landed

    const int tid = threadIdx.x;
    const int pair = tid / 20, col = tid - pair * 20;
    const uint32_t bar = smem_addr(&s_bar);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1u));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t bulk_phase = 0;                                   // parity of the next bulk completion
    bool on_bar = false;                                       // the tile at the top of the loop arrives on the mbarrier
    if ((int)blockIdx.x < total_tiles) on_bar = stage_pcm(s_pcm, bar, pcm, tiles[blockIdx.x], tid);
    for (int i = tid; i < 400; i += MEL_THREADS) {
        s_win[i] = tab->window[i];
        const int k = ((i / 20) * (i % 20)) % 400;
        s_twr[i] = tab->tw_re[k];
        s_twi[i] = tab->tw_im[k];
    }
    for (int i = tid; i < FBW_MAX; i += MEL_THREADS) s_fbw[i] = 0.25f * tab->fb_w[i];
    for (int i = tid; i < 3 * NM; i += MEL_THREADS) s_fbi[i] = tab->fb_idx[i];
    if (tid == 0) s_tmax = (int)0xff800000u;                   // -inf

    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const MelTile cur = tiles[tile];                             // used from the mel phase on: the load hides under step 1
        if (on_bar) {
            int spins = 0;
            while (!mbar_try(bar, bulk_phase)) if (++spins > 20000) __trap();      // bounded: a bug must not hang the GPU
            bulk_phase ^= 1u;
        } else {
            asm volatile("cp.async.wait_all;\n" ::: "memory");
        }
        __syncthreads();                                             // this tile's PCM is in; the previous tile's output is out

        float2* scr = s_scr + pair * SCR;
        // ---- step 1: 20-pt DFT over n1 for column n2 = col, twiddle, scatter ----
        {
            const float* pa = s_pcm + pair * (320 + SKEW) + col;     // sample i of frame A: pa[i + SKEW*(i >= 320)]
            c32 v[20];
#pragma unroll
            for (int n1 = 0; n1 < 20; ++n1) {
                const float w = s_win[20 * n1 + col];
                v[n1] = {pa[20 * n1 + (n1 >= 16 ? SKEW : 0)] * w, pa[160 + 20 * n1 + (n1 >= 8 ? SKEW : 0)] * w};
            }
            if (PACKED) fft_packed::dft20(v); else fft_scalar::dft20(v);
#pragma unroll
            for (int k1 = 0; k1 < 20; ++k1) {
                c32 z = fft_scalar::cmul(v[k1], c32{s_twr[k1 * 20 + col], s_twi[k1 * 20 + col]});
                scr[k1 * 21 + col] = make_float2(z.x, z.y);
            }
        }
        __syncthreads();
        {   // the PCM buffer is free: start the next tile's copy under the rest of this one
            const int next = tile + gridDim.x;
            on_bar = next < total_tiles && stage_pcm(s_pcm, bar, pcm, tiles[next], tid);
        }
        // ---- step 2: 20-pt DFT over n2 for row k1 = col; the spectrum of pair p lands at scr_p + (p >> 2), so that
        //      the twelve pairs sit in twelve different bank pairs for the pair-parallel reads below ----
        {
            c32 v[20];
#pragma unroll
            for (int n2 = 0; n2 < 20; ++n2) {
                float2 t = scr[col * 21 + n2];
                v[n2] = {t.x, t.y};
            }
            if (PACKED) fft_packed::dft20(v); else fft_scalar::dft20(v);
            __syncthreads();                                         // every row has been read
            float2* zp = scr + (pair >> 2) + col;
#pragma unroll
            for (int k2 = 0; k2 < 20; ++k2) zp[20 * k2] = make_float2(v[k2].x, v[k2].y);
        }
        __syncthreads();
        // ---- power spectra of both packed frames, k = 0..200: 2A = Z[k] + conj Z[400-k], 2iB = Z[k] - conj Z[400-k]
        //      (through registers: the rows overwrite the scratch; the 1/4 is folded into the filterbank weights) ----
        {
            constexpr int NIT = (NPAIR * 201 + MEL_THREADS - 1) / MEL_THREADS;
            float pa[NIT], pb[NIT];
#pragma unroll
            for (int r = 0; r < NIT; ++r) {
                const int it = tid + r * MEL_THREADS;
                if (it < NPAIR * 201) {
                    const int p = it / 201, k = it - p * 201;
                    const float2* Z = s_scr + p * SCR + (p >> 2);
                    const float2 z = Z[k];
                    const float2 c = Z[k == 0 ? 0 : 400 - k];
                    const float ar = z.x + c.x, ai = z.y - c.y;
                    const float br = z.x - c.x, bi = z.y + c.y;
                    pa[r] = fmaf(ar, ar, ai * ai);
                    pb[r] = fmaf(br, br, bi * bi);
                }
            }
            __syncthreads();
#pragma unroll
            for (int r = 0; r < NIT; ++r) {
                const int it = tid + r * MEL_THREADS;
                if (it < NPAIR * 201) {
                    const int p = it / 201, k = it - p * 201;
                    s_pow[tile_row(2 * p) * PPL + k] = pa[r];
                    s_pow[tile_row(2 * p + 1) * PPL + k] = pb[r];
                }
            }
        }
        __syncthreads();
        // ---- mel filterbank (f32 sum in k order, main.rs:484-490) for 4 frames at a time, log10 -> s_out[frame][OP] ----
        float tmax = -INFINITY;
        for (int it = tid; it < NM * (FPT / 4); it += MEL_THREADS) {
            const int m = it / (FPT / 4), fg = it - m * (FPT / 4);
            const float* p = s_pow + fg * PPL + s_fbi[m];              // rows fg, fg + 6, fg + 12, fg + 18 = frames 4 fg + q
            const float* w = s_fbw + s_fbi[2 * NM + m];
            const int len = s_fbi[NM + m];
            float e0 = 0.0f, e1 = 0.0f, e2 = 0.0f, e3 = 0.0f;
            for (int j = 0; j < len; ++j) {
                const float wj = w[j];
                e0 = fmaf(wj, p[j], e0);
                e1 = fmaf(wj, p[6 * PPL + j], e1);
                e2 = fmaf(wj, p[12 * PPL + j], e2);
                e3 = fmaf(wj, p[18 * PPL + j], e3);
            }
            const float e[4] = {e0, e1, e2, e3};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float lv = 0.30102999566398120f * __log2f(fmaxf(e[q], 1e-10f));
                s_out[(fg + 6 * q) * OP + m] = lv;
                if (cur.f0 + 4 * fg + q < cur.nf) tmax = fmaxf(tmax, lv);
            }
        }
        {   // 15 half-warps: reduce inside each (the CTA's last warp has only its lower half)
            const unsigned mask = tid >= (MEL_THREADS & ~31) ? 0x0000ffffu : 0xffffffffu;
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) tmax = fmaxf(tmax, __shfl_xor_sync(mask, tmax, o, 16));
            if ((tid & 15) == 0 && tmax > -INFINITY) atomic_max_float(&s_tmax, tmax);
        }
        __syncthreads();
        {
            const int left = cur.nf - cur.f0;
            const int n_valid = (left < FPT ? left : FPT) * NM;
            float* dst = raw + cur.raw_base * NM;
            for (int j = tid; j < n_valid; j += MEL_THREADS) {
                const int f = j / NM;
                dst[j] = s_out[tile_row(f) * OP + (j - f * NM)];
            }
        }
        if (tid == 0) {
            const int m = s_tmax;
            if (m != (int)0xff800000u) atomic_max_float(fmax + cur.file, __int_as_float(m));
            s_tmax = (int)0xff800000u;
        }
        // the barrier at the top of the next tile (after its PCM wait) frees s_out / s_pow / s_tmax
    }
}

__global__ void fill_int_kernel(int* __restrict__ p, int n, int v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

template <typename T> __device__ __forceinline__ T to_out(float v);
template <> __device__ __forceinline__ float to_out<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 to_out<__nv_bfloat16>(float v) { return __float2bfloat16(v); }

__device__ __forceinline__ void store4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void store4(__nv_bfloat16* p, float4 v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&a);
    u.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = u;
}

// K1b: grid (x, chunk); out [chunk][3002][NM], rows 0 and 3001 stay zero.  Four mel bins per thread (128-bit loads).
template <typename T, int NM>
__global__ void mel_chunks_kernel(const float* __restrict__ raw, const int64_t* __restrict__ frame_off,
                                  const int* __restrict__ fmax, const MelChunk* __restrict__ chunks,
                                  T* __restrict__ out, int chunk0) {
    const int c = blockIdx.y;
    const MelChunk ch = chunks[chunk0 + c];
    const int64_t nf = frame_off[ch.file + 1] - frame_off[ch.file];
    const float floor_v = __int_as_float(fmax[ch.file]) - 8.0f;
    const float4* src = reinterpret_cast<const float4*>(raw + (frame_off[ch.file] + ch.frame_start) * NM);
    T* dst = out + (size_t)c * (WB_N_FRAMES + 2) * NM;
    constexpr int V = NM / 4;
    const int total = (WB_N_FRAMES + 2) * V;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int t = i / V - 1;
        float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);                // literal 0.0 padding (Q2)
        if (t >= 0 && t < WB_N_FRAMES && ch.frame_start + t < nf) {
            const float4 lv = __ldg(src + (i - V));
            v.x = (fmaxf(lv.x, floor_v) + 4.0f) / 4.0f;               // main.rs:502-506
            v.y = (fmaxf(lv.y, floor_v) + 4.0f) / 4.0f;
            v.z = (fmaxf(lv.z, floor_v) + 4.0f) / 4.0f;
            v.w = (fmaxf(lv.w, floor_v) + 4.0f) / 4.0f;
        }
        store4(dst + (size_t)i * 4, v);
    }
}

// K1c: reference layout out[m][f] for f in [0,n_out), source frames frame0+f (valid if < nf).
template <int NM>
__global__ void mel_export_kernel(const float* __restrict__ raw, int64_t frame_base, int64_t frame0,
                                  int64_t nf, const int* __restrict__ fmax, int file,
                                  float* __restrict__ out, int64_t n_out) {
    __shared__ float tile[32][NM + 1];
    const float floor_v = __int_as_float(fmax[file]) - 8.0f;
    const int64_t fb = (int64_t)blockIdx.x * 32;
    for (int i = threadIdx.x; i < 32 * NM; i += blockDim.x) {
        int fl = i / NM, m = i - fl * NM;
        int64_t f = frame0 + fb + fl;
        float v = 0.0f;
        if (fb + fl < n_out && f < nf) v = (fmaxf(raw[(frame_base + f) * NM + m], floor_v) + 4.0f) / 4.0f;
        tile[fl][m] = v;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * NM; i += blockDim.x) {
        int m = i / 32, fl = i - m * 32;
        if (fb + fl < n_out) out[(int64_t)m * n_out + fb + fl] = tile[fl][m];
    }
}

// host mel [B][n_mels][3000] (reference layout) -> time-major padded compute-dtype buffer
template <typename T>
__global__ void mel_transpose_in_kernel(const float* __restrict__ in, T* __restrict__ out, int n_mels) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int t0 = blockIdx.x * 32, m0 = blockIdx.y * 32;
    const float* src = in + (size_t)b * n_mels * WB_N_FRAMES;
    T* dst = out + (size_t)b * (WB_N_FRAMES + 2) * n_mels;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int m = m0 + i, t = t0 + threadIdx.x;
        tile[i][threadIdx.x] = (m < n_mels && t < WB_N_FRAMES) ? src[(size_t)m * WB_N_FRAMES + t] : 0.0f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int t = t0 + i, m = m0 + threadIdx.x;
        if (t < WB_N_FRAMES && m < n_mels) dst[(size_t)(t + 1) * n_mels + m] = to_out<T>(tile[threadIdx.x][i]);
    }
}

constexpr size_t mel_smem(int nm) { return sizeof(float) * (size_t)(PCM_LD + 400 + 800 + FBW_MAX + 3 * nm + ((3 * nm) & 1) + 2 * NPAIR * SCR); }

}  // namespace

// ---------------- host side ----------------
void mel_build_tables(MelTables& t, int n_mels) {
    if (n_mels != 80 && n_mels != 128) return;      // encoder-only use with host mel input; wb_upload_pcm rejects the config
    // main.rs:323-330
    for (int i = 0; i < 400; ++i) {
        float x = (3.14159265358979323846f * 2.0f * (float)i) / 400.0f;
        t.window[i] = 0.5f - 0.5f * cosf(x);
    }
    for (int k = 0; k < 400; ++k) {
        double a = -2.0 * 3.14159265358979323846 * (double)k / 400.0;
        t.tw_re[k] = (float)cos(a);
        t.tw_im[k] = (float)sin(a);
    }
    // main.rs:332-405 (Slaney mel scale + area normalisation), f32 throughout
    auto hz_to_mel = [](float hz) {
        const float logstep = 27.0f / logf(6.4f);
        float mel = 3.0f * hz / 200.0f;
        if (hz >= 1000.0f) mel = 15.0f + logf(hz / 1000.0f) * logstep;
        return mel;
    };
    auto mel_to_hz = [](float mel) {
        const float logstep = logf(6.4f) / 27.0f;
        float hz = 200.0f * mel / 3.0f;
        if (mel >= 15.0f) hz = 1000.0f * expf(logstep * (mel - 15.0f));
        return hz;
    };
    const int n_freq = 201;
    float fmax_hz = fminf(8000.0f, 16000.0f / 2.0f);
    float mel_min = hz_to_mel(0.0f), mel_max = hz_to_mel(fmax_hz);
    float fp[130], ff[201];
    for (int i = 0; i < n_mels + 2; ++i) {
        float m = mel_min + (mel_max - mel_min) * (float)i / (float)(n_mels + 1);
        fp[i] = mel_to_hz(m);
    }
    for (int k = 0; k < n_freq; ++k) ff[k] = (float)k * 8000.0f / (float)(n_freq - 1);
    std::vector<float> fb((size_t)n_mels * n_freq);
    for (int m = 0; m < n_mels; ++m) {
        float dl = fmaxf(fp[m + 1] - fp[m], 1e-6f), dr = fmaxf(fp[m + 2] - fp[m + 1], 1e-6f);
        float enorm = 2.0f / fmaxf(fp[m + 2] - fp[m], 1e-6f);
        for (int k = 0; k < n_freq; ++k) {
            float lower = (ff[k] - fp[m]) / dl, upper = (fp[m + 2] - ff[k]) / dr;
            float w = fmaxf(fminf(lower, upper), 0.0f);
            fb[(size_t)m * n_freq + k] = w * enorm;
        }
    }
    // compact: contiguous non-zero run per mel (zeros contribute exactly 0 to the f32 sum)
    int off = 0;
    for (int i = 0; i < FBW_MAX; ++i) t.fb_w[i] = 0.0f;
    for (int m = 0; m < n_mels; ++m) {
        int s = 0, e = 0;
        bool any = false;
        for (int k = 0; k < n_freq; ++k)
            if (fb[(size_t)m * n_freq + k] != 0.0f) {
                if (!any) s = k;
                any = true;
                e = k + 1;
            }
        if (!any) { s = 0; e = 0; }
        WB_REQUIRE(off + (e - s) <= FBW_MAX, WB_EINVAL, "mel filterbank has too many non-zeros");
        t.fb_idx[m] = s;
        t.fb_idx[n_mels + m] = e - s;
        t.fb_idx[2 * n_mels + m] = off;
        for (int k = s; k < e; ++k) t.fb_w[off++] = fb[(size_t)m * n_freq + k];
    }
}

int64_t mel_n_frames(int64_t n) {
    int64_t nf = 1 + n / 160;          // 1 + (n + 400 - 400)/160, main.rs:444-448
    if (nf > 1) nf -= 1;               // dropped last frame, :450-452
    return nf;
}

// cudaFuncSetAttribute applies to the CURRENT device only: called from wb_create for every context (after
// cudaSetDevice), never behind a process-wide flag (a second GPU in the same process would miss the opt-in).
void mel_set_attrs() {
    CUDA_CHECK(cudaFuncSetAttribute(logmel_raw_kernel<80, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mel_smem(80)));
    CUDA_CHECK(cudaFuncSetAttribute(logmel_raw_kernel<80, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mel_smem(80)));
    CUDA_CHECK(cudaFuncSetAttribute(logmel_raw_kernel<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mel_smem(128)));
    CUDA_CHECK(cudaFuncSetAttribute(logmel_raw_kernel<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mel_smem(128)));
}

void mel_build_tiles(wb_ctx* ctx) {
    MelState& s = ctx->mel;
    s.tiles.reserve((size_t)s.total_tiles);
    mel_tiles_kernel<<<ceil_div(s.total_tiles, 256), 256, 0, ctx->stream>>>(s.file_off.p, s.frame_off.p, s.tile_off.p, s.n_files_staged,
                                                                            s.total_tiles, s.tiles.p);
    CUDA_CHECK(cudaGetLastError());
}

void mel_launch_raw(wb_ctx* ctx) {
    MelState& s = ctx->mel;
    fill_int_kernel<<<ceil_div(s.n_files, 256), 256, 0, ctx->stream>>>(s.fmax.p, s.n_files, (int)0xff800000u);   // -inf as int bits
    int per_sm = 3;                                         // resident CTAs per SM (registers and shared memory both allow 3)
    if (const char* e = getenv("WB_MEL_CTAS_PER_SM")) per_sm = std::max(1, atoi(e));
    const int grid = std::min(s.total_tiles, per_sm * ctx->sm_count);
    bool packed = true;                                     // FADD2 butterflies
    if (const char* e = getenv("WB_MEL_PACKED")) packed = atoi(e) != 0;
#define WB_MEL_RAW(NM, PK) logmel_raw_kernel<NM, PK><<<grid, MEL_THREADS, mel_smem(NM), ctx->stream>>>( \
        s.pcm.p, s.tiles.p, s.total_tiles, ctx->mel_tables_dev, s.raw.p, s.fmax.p)
    if (ctx->cfg.n_mels == 128) { if (packed) WB_MEL_RAW(128, true); else WB_MEL_RAW(128, false); }
    else { if (packed) WB_MEL_RAW(80, true); else WB_MEL_RAW(80, false); }
#undef WB_MEL_RAW
    CUDA_CHECK(cudaGetLastError());
    ctx->timing.mel_launches += 2;          // fill + K1a
}

void mel_launch_chunks(wb_ctx* ctx, int chunk0, int n, void* out) {
    MelState& s = ctx->mel;
    dim3 grid(30, n);
    const bool bf = ctx->cfg.precision == WB_PREC_BF16, big = ctx->cfg.n_mels == 128;
#define WB_MEL_CHUNKS(T, NM) mel_chunks_kernel<T, NM><<<grid, 256, 0, ctx->stream>>>(s.raw.p, s.frame_off.p, s.fmax.p, s.chunks.p, (T*)out, chunk0)
    if (bf && big) WB_MEL_CHUNKS(__nv_bfloat16, 128);
    else if (bf) WB_MEL_CHUNKS(__nv_bfloat16, 80);
    else if (big) WB_MEL_CHUNKS(float, 128);
    else WB_MEL_CHUNKS(float, 80);
#undef WB_MEL_CHUNKS
    CUDA_CHECK(cudaGetLastError());
    ctx->timing.mel_launches += 1;
}

void mel_launch_export(wb_ctx* ctx, int file, int64_t frame0, int64_t n_out, float* out_dev) {
    MelState& s = ctx->mel;
    int64_t nf = s.h_frame_off[file + 1] - s.h_frame_off[file];
    int blocks = (int)ceil_div64(n_out, 32);
    if (ctx->cfg.n_mels == 128)
        mel_export_kernel<128><<<blocks, 256, 0, ctx->stream>>>(s.raw.p, s.h_frame_off[file], frame0, nf, s.fmax.p, file, out_dev, n_out);
    else
        mel_export_kernel<80><<<blocks, 256, 0, ctx->stream>>>(s.raw.p, s.h_frame_off[file], frame0, nf, s.fmax.p, file, out_dev, n_out);
    CUDA_CHECK(cudaGetLastError());
}

void mel_launch_transpose_in(wb_ctx* ctx, const float* in_dev, void* out, int B) {
    int n_mels = ctx->cfg.n_mels;
    dim3 grid(ceil_div(WB_N_FRAMES, 32), ceil_div(n_mels, 32), B), block(32, 8);
    if (ctx->cfg.precision == WB_PREC_BF16)
        mel_transpose_in_kernel<__nv_bfloat16><<<grid, block, 0, ctx->stream>>>(in_dev, (__nv_bfloat16*)out, n_mels);
    else
        mel_transpose_in_kernel<float><<<grid, block, 0, ctx->stream>>>(in_dev, (float*)out, n_mels);
    CUDA_CHECK(cudaGetLastError());
}

// CPU emulation of the kernel's FFT data flow (same mel_math.h code) for the no-GPU unit test.
extern "C" int wb_selftest_fft400(const float* re, const float* im, float* out_re, float* out_im) {
    using namespace fft_scalar;
    std::vector<c32> scr(420), X(400);
    for (int n2 = 0; n2 < 20; ++n2) {
        c32 v[20];
        for (int n1 = 0; n1 < 20; ++n1) v[n1] = {re[20 * n1 + n2], im[20 * n1 + n2]};
        dft20(v);
        for (int k1 = 0; k1 < 20; ++k1) {
            double a = -2.0 * 3.14159265358979323846 * (double)(n2 * k1) / 400.0;
            scr[k1 * 21 + n2] = cmul(v[k1], c32{(float)cos(a), (float)sin(a)});
        }
    }
    for (int k1 = 0; k1 < 20; ++k1) {
        c32 v[20];
        for (int n2 = 0; n2 < 20; ++n2) v[n2] = scr[k1 * 21 + n2];
        dft20(v);
        for (int k2 = 0; k2 < 20; ++k2) X[k1 + 20 * k2] = v[k2];
    }
    for (int k = 0; k < 400; ++k) { out_re[k] = X[k].x; out_im[k] = X[k].y; }
    return 0;
}
