// vocab_tc.cu — the vocabulary projection of a decode step on tcgen05 (kernel K3g of the bf16 build), with the
// final LayerNorm in front and the masked arg-max behind it:
//     logits[n][b] = E[n][:] . LN(x[b][:])          n < vocab (51865), b < 32 sequences, K = d_model
// in the swap-AB view the round-1 verdict asked for: the 128 rows of a weight tile are the M dimension of one
// tcgen05.mma (cta_group::1, kind::f16, 128 x 32 x 16), the <= 32 sequences are N, the accumulator is 32 TMEM columns.
// Weight tiles [128 rows][64 k] arrive by TMA (128-byte swizzle, rows past the vocabulary zero-filled), two to a stage,
// through a ring of 5 stages (160 KB in flight per SM) and do not depend on the predecessor kernel, so the first five
// stages are requested BEFORE the programmatic-dependent-launch wait; the activations are LayerNorm-ed once per CTA into the swizzled K-major
// layout the B operand wants.  Four epilogue warps read the accumulator with tcgen05.ld (thread = weight row, registers =
// sequences) and keep a running masked arg-max per sequence (strict '>', lowest index wins ties, NaN never wins:
// argmax_last_dim_raw, /root/reference/src/main.rs:709-735); one partial per CTA and sequence goes to
// argmax_merge_kernel (decoder.cu).  No logits are written: this path serves decodes that did not ask for them.
// Replaces skinny_mma_kernel<8,1,16> (legacy mma.sync: tensor-issue-bound at 3.7 TB/s even with the weights in L2).
#include <cuda.h>

#include <utility>
#include <vector>

#include "ctx.h"

namespace {

constexpr int VT_THREADS = 192;                       // warp 0: TMA producer, warp 1: MMA issue + TMEM, warps 2-5: arg-max epilogue
constexpr int VT_BM = 128, VT_BK = 64, VT_N = 32;
constexpr int VT_STAGES = 5;                                    // stages of up to 2 k-blocks (32 KB): 160 KB of weights in flight per SM and
constexpr int VT_KPS_MAX = 2;                                   //   one mbarrier round trip per 32 KB.  Deliberately NOT the whole 227 KB:
                                                                //   with 193 KB a cross-attention CTA of another batch in flight still fits
                                                                //   on the SM (3 x 64 KB stages: 16.5 -> same alone, 40.6 -> 37.9 k in flight)
constexpr uint32_t VT_KB_BYTES = VT_BM * VT_BK * 2;             // one [128 rows][64 k] k-block: 16 KB
constexpr uint32_t VT_STAGE_BYTES = VT_KPS_MAX * VT_KB_BYTES;   // 64 KB
constexpr uint32_t VT_BTILE_BYTES = VT_N * VT_BK * 2;           // 4 KB: [32 sequences][64 k]
constexpr int VT_MAX_KB = 8;                                    // d_model <= 512
constexpr uint32_t VT_TMEM_COLS = 64;                           // two accumulators of 32 columns

struct VocabTc {
    CUtensorMap tmW;
    bool ok = false;
    bool skinny_ok = false;
    std::vector<std::pair<const void*, CUtensorMap>> linear_maps;      // one per decoder linear weight ([out][in] bf16)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Bounded wait: a pipeline bug must end in a trap (launch error), never in a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    for (int spins = 0; !done; ++spins) {
        asm volatile(
            "{\n.reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 100000;\n"
            "selp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (!done && spins > 20000) __trap();
    }
}
// the same load with an L2 eviction-priority hint (the vocabulary matrix is re-read by every step of every batch in flight)
__device__ __forceinline__ void tma_load_2d_hint(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(pol)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor: rows of 128 bytes, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// kind::f16 instruction descriptor: D = f32, A = B = bf16, both K-major, M = 128, N = 32
__device__ __forceinline__ constexpr uint32_t make_idesc(int n = VT_N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(VT_BM >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}

// NG = groups of 32 sequences (1 or 2): the tcgen05.mma is 128 x (32 NG) x 16, the accumulators 2 x 32 NG TMEM columns.
// With 64 sequences the ring has 4 stages (128 KB of weights in flight) next to 64 KB of staged activations.
template <int NG> struct VtCfg {
    static constexpr int N = VT_N * NG;                                    // sequences = N of the MMA
    static constexpr int STAGES = NG == 1 ? VT_STAGES : 4;
    static constexpr uint32_t BTILE = VT_BTILE_BYTES * NG;                // [N sequences][64 k]
    static constexpr uint32_t TMEM_COLS = 2 * N;
    static constexpr size_t SMEM = (size_t)STAGES * VT_STAGE_BYTES + VT_MAX_KB * BTILE + 1024 /*align*/ + 512 /*barriers*/;
};

template <int NG>
__global__ void __launch_bounds__(VT_THREADS, 1)
vocab_tc_kernel(const __grid_constant__ CUtensorMap tmW, const float* __restrict__ X, int B, int K, int N,
                const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                const int* __restrict__ state, const unsigned* __restrict__ sup_base, const unsigned* __restrict__ sup_first,
                float* __restrict__ amax_val, int* __restrict__ amax_idx, int l2_last) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    constexpr int NSEQ = VtCfg<NG>::N, STAGES = VtCfg<NG>::STAGES;
    constexpr uint32_t BTILE = VtCfg<NG>::BTILE;
    uint8_t* sA = smem;                                                   // STAGES weight tiles
    uint8_t* sB = smem + STAGES * VT_STAGE_BYTES;                         // K/64 activation tiles
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + VT_MAX_KB * BTILE);
    uint64_t* full = bars;                     // [STAGES <= VT_STAGES] weight tile landed
    uint64_t* empty = bars + VT_STAGES;        // [STAGES] MMAs that read the stage retired
    uint64_t* acc_full = bars + 2 * VT_STAGES;     // [2] accumulator complete
    uint64_t* acc_empty = bars + 2 * VT_STAGES + 2; // [2] accumulator drained by the 128 epilogue threads
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * VT_STAGES + 4);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nkb = K / VT_BK;
    const int kps = nkb % VT_KPS_MAX == 0 ? VT_KPS_MAX : 1;                             // k-blocks per stage
    const int gpt = nkb / kps;                                                          // stages per tile
    const int n_tiles = (N + VT_BM - 1) / VT_BM;
    const int my_tiles = (int)blockIdx.x < n_tiles ? (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int total_g = my_tiles * gpt;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmW)) : "memory");
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(&acc_full[0], 1); mbar_init(&acc_full[1], 1);
        mbar_init(&acc_empty[0], 128); mbar_init(&acc_empty[1], 128);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(VtCfg<NG>::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    uint64_t wpol;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(wpol));
    auto issue_load = [&](int g) {                                        // g-th stage of this CTA's tile sequence
        const int ti = g / gpt, grp = g - ti * gpt;
        const int tile = (int)blockIdx.x + ti * (int)gridDim.x;
        const int s = g % STAGES;
        mbar_expect_tx(&full[s], (uint32_t)kps * VT_KB_BYTES);
        for (int j = 0; j < kps; ++j) {
            if (l2_last) tma_load_2d_hint(&tmW, &full[s], sA + (size_t)s * VT_STAGE_BYTES + (size_t)j * VT_KB_BYTES, (grp * kps + j) * VT_BK, tile * VT_BM, wpol);
            else tma_load_2d(&tmW, &full[s], sA + (size_t)s * VT_STAGE_BYTES + (size_t)j * VT_KB_BYTES, (grp * kps + j) * VT_BK, tile * VT_BM);
        }
    };
    // weights are constant during a decode: the first ring-full leaves before the predecessor kernel has finished
    int issued = 0;
    if (warp == 0 && lane == 0) {
        const int first = total_g < STAGES ? total_g : STAGES;
        for (; issued < first; ++issued) issue_load(issued);
    }
    // LayerNorm parameters are weights too: fetch them under the predecessor's tail
    float4 gw[4], gb[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = i * 128 + lane * 4;
        gw[i] = c < K ? __ldg(reinterpret_cast<const float4*>(ln_w + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
        gb[i] = c < K ? __ldg(reinterpret_cast<const float4*>(ln_b + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    // ---- LayerNorm of the <= 32 sequences -> bf16, K-major 128B-swizzled tiles [kb][32 rows][64 k] (the B operand) ----
    const unsigned* sup = (state[0] - (state[1] - 1) == 0) ? sup_first : sup_base;     // first generated token?
    for (int grp0 = 0; grp0 < NSEQ; grp0 += VT_N) {                        // 32 sequences at a time
        // warp w owns rows w, w+6, ..: all their loads leave together (one memory round trip for the whole prologue)
        constexpr int NW = VT_THREADS / 32, RPW = (VT_N + NW - 1) / NW;   // 6 warps, <= 6 rows each
        float4 xv[RPW][4];
        float s1[RPW], qq[RPW];
#pragma unroll
        for (int j = 0; j < RPW; ++j) {
            const int r = grp0 + warp + j * NW;
            s1[j] = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int c = i * 128 + lane * 4;
                xv[j][i] = (warp + j * NW < VT_N && r < B && c < K) ? __ldcg(reinterpret_cast<const float4*>(X + (size_t)r * K + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int j = 0; j < RPW; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) s1[j] += (xv[j][i].x + xv[j][i].y) + (xv[j][i].z + xv[j][i].w);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int j = 0; j < RPW; ++j) s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], o);
#pragma unroll
        for (int j = 0; j < RPW; ++j) {
            const float mean = s1[j] / (float)K;
            qq[j] = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (i * 128 + lane * 4 < K) {
                    const float t0 = xv[j][i].x - mean, t1 = xv[j][i].y - mean, t2 = xv[j][i].z - mean, t3 = xv[j][i].w - mean;
                    qq[j] += (t0 * t0 + t1 * t1) + (t2 * t2 + t3 * t3);
                }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int j = 0; j < RPW; ++j) qq[j] += __shfl_xor_sync(0xffffffffu, qq[j], o);
#pragma unroll
        for (int j = 0; j < RPW; ++j) {
            if (warp + j * NW >= VT_N) continue;
            const int r = grp0 + warp + j * NW;
            const float mean = s1[j] / (float)K;
            const float rs = 1.0f / sqrtf(qq[j] / (float)K + 1e-5f);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int c = i * 128 + lane * 4;
                if (c < K) {
                    uint2 pk = make_uint2(0u, 0u);
                    if (r < B) {
                        pk.x = pack_bf16x2((xv[j][i].x - mean) * rs * gw[i].x + gb[i].x, (xv[j][i].y - mean) * rs * gw[i].y + gb[i].y);
                        pk.y = pack_bf16x2((xv[j][i].z - mean) * rs * gw[i].z + gb[i].z, (xv[j][i].w - mean) * rs * gw[i].w + gb[i].w);
                    }
                    const int kb = c >> 6, chunk = (c & 63) >> 3;         // 16-byte chunk of the 128-byte row
                    *reinterpret_cast<uint2*>(sB + (size_t)kb * BTILE + r * 128 + ((chunk ^ (r & 7)) << 4) + (c & 7) * 2) = pk;
                }
            }
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // generic-proxy stores -> visible to the tensor core
    __syncthreads();

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer: keep the ring full =====
            for (; issued < total_g; ++issued) {
                const int s = issued % STAGES;
                mbar_wait(&empty[s], (uint32_t)((issued / STAGES - 1) & 1));
                issue_load(issued);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issue =====
            const uint32_t idesc = make_idesc(NSEQ);
            int g = 0;
            for (int ti = 0; ti < my_tiles; ++ti) {
                const int acc = ti & 1;
                if (ti >= 2) mbar_wait(&acc_empty[acc], (uint32_t)(((ti >> 1) - 1) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int grp = 0; grp < gpt; ++grp, ++g) {
                    const int s = g % STAGES;
                    mbar_wait(&full[s], (uint32_t)((g / STAGES) & 1));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    for (int j = 0; j < kps; ++j) {
                        const int kb = grp * kps + j;
                        const uint64_t da = make_smem_desc(smem_u32(sA + (size_t)s * VT_STAGE_BYTES + (size_t)j * VT_KB_BYTES));
                        const uint64_t db = make_smem_desc(smem_u32(sB + (size_t)kb * BTILE));
#pragma unroll
                        for (int k = 0; k < VT_BK / 16; ++k)
                            umma(tmem_base + acc * NSEQ, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (uint32_t)((kb | k) != 0));
                    }
                    umma_commit(&empty[s]);                               // the stage is free once these MMAs retire
                }
                umma_commit(&acc_full[acc]);
            }
        }
    } else {
        // ===== epilogue: thread = weight row of the tile (TMEM lane), registers = the 32 sequences =====
        const int q = warp & 3;                                           // TMEM lane quadrant this warp may read
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        float bestv[NSEQ];
        int besti[NSEQ];
#pragma unroll
        for (int s = 0; s < NSEQ; ++s) { bestv[s] = -INFINITY; besti[s] = 0x7fffffff; }
        for (int ti = 0; ti < my_tiles; ++ti) {
            const int acc = ti & 1;
            mbar_wait(&acc_full[acc], (uint32_t)((ti >> 1) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int n = ((int)blockIdx.x + ti * (int)gridDim.x) * VT_BM + q * 32 + lane;
            const bool ok = n < N && !((sup[n >> 5] >> (n & 31)) & 1u);
#pragma unroll
            for (int gI = 0; gI < NG; ++gI) {
                uint32_t r[32];
                tmem_ld32(tmem_base + lane_off + acc * NSEQ + gI * VT_N, r);
                if (gI == NG - 1) {
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    mbar_arrive(&acc_empty[acc]);
                }
                if (ok) {
#pragma unroll
                    for (int s = 0; s < VT_N; ++s) {
                        const float v = __uint_as_float(r[s]);
                        if (v > bestv[gI * VT_N + s]) { bestv[gI * VT_N + s] = v; besti[gI * VT_N + s] = n; }   // rows only grow with ti: strict '>' keeps the lowest index
                    }
                }
            }
        }
        // ---- 128 rows -> one partial per sequence: through shared memory (the weight ring is idle by now) ----
        float* sv = reinterpret_cast<float*>(sA);                         // [128][33]
        int* si = reinterpret_cast<int*>(sA + 128 * 33 * 4);
        float* pv = reinterpret_cast<float*>(sA + 2 * 128 * 33 * 4);      // [4][32] values | indices
        int* pi = reinterpret_cast<int*>(pv + 4 * 32);
        const int row = q * 32 + lane;
        // every MMA that reads sA has retired (the last acc_full was waited for above); TMA wrote only what they read
#pragma unroll
        for (int gI = 0; gI < NG; ++gI) {
            if (gI > 0) asm volatile("bar.sync 1, 128;" ::: "memory");    // the previous group's scans and merge are done
#pragma unroll
            for (int s = 0; s < VT_N; ++s) { sv[row * 33 + s] = bestv[gI * VT_N + s]; si[row * 33 + s] = besti[gI * VT_N + s]; }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            // thread (sequence = lane, quarter = q) scans 32 of the 128 rows; the four quarters meet in shared memory
            float bv = -INFINITY;
            int bi = 0x7fffffff;
#pragma unroll 8
            for (int rr = q * 32; rr < q * 32 + 32; ++rr) {
                const float v = sv[rr * 33 + lane];
                const int i = si[rr * 33 + lane];
                if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
            }
            pv[q * 32 + lane] = bv;
            pi[q * 32 + lane] = bi;
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (warp == 2) {
                bv = pv[lane]; bi = pi[lane];
#pragma unroll
                for (int k = 1; k < 4; ++k) {
                    const float v = pv[k * 32 + lane];
                    const int i = pi[k * 32 + lane];
                    if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
                }
                amax_val[(size_t)blockIdx.x * NSEQ + gI * VT_N + lane] = bv;
                amax_idx[(size_t)blockIdx.x * NSEQ + gI * VT_N + lane] = bi;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(VtCfg<NG>::TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// The same machine for the per-layer GEMMs of a decode step (q|k|v, out, cross-q, cross-out, fc1, fc2):
//     Y[b][n] = epi( W[n][:] . f(X[b][:]) ),   f = LayerNorm or identity,   epi = + bias, GELU, + residual
// One CTA per 128 weight rows (4 .. 16 CTAs per GEMM instead of 16 .. 32 register-heavy mma.sync CTAs), weights through
// the TMA ring (no registers), accumulator [128 rows][32 sequences] in TMEM, epilogue thread = weight row: its 32 values
// go to Y[b][n] with the lanes of a warp on consecutive n (coalesced).  Replaces skinny_mma_kernel for d_model <= 512.
struct SkinnyTcArgs {
    const float* X; int B, K, N;
    const float* ln_w; const float* ln_b; const float* bias;
    int act;
    const float* residual;
    float* Y;
    int n_stages;
};

__device__ __forceinline__ float gelu_erf_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

__global__ void __launch_bounds__(VT_THREADS, 1)
skinny_tc_kernel(const __grid_constant__ CUtensorMap tmW, const SkinnyTcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    const int K = a.K, N = a.N, B = a.B;
    const int nkb = K / VT_BK;
    uint8_t* sB = smem;                                                   // K/64 activation tiles of 4 KB
    uint8_t* sA = smem + (size_t)nkb * VT_BTILE_BYTES;                    // n_stages weight stages of 32 KB
    uint64_t* bars = reinterpret_cast<uint64_t*>(sA + (size_t)a.n_stages * VT_STAGE_BYTES);
    uint64_t* full = bars;                     // [n_stages <= VT_STAGES]
    uint64_t* empty = bars + VT_STAGES;
    uint64_t* acc_full = bars + 2 * VT_STAGES;
    uint64_t* acc_empty = bars + 2 * VT_STAGES + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * VT_STAGES + 4);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int kps = nkb % VT_KPS_MAX == 0 ? VT_KPS_MAX : 1;
    const int gpt = nkb / kps;
    const int n_tiles = N / VT_BM;
    const int my_tiles = (int)blockIdx.x < n_tiles ? (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int total_g = my_tiles * gpt;
    const int NS = a.n_stages;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmW)) : "memory");
        for (int i = 0; i < NS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(&acc_full[0], 1); mbar_init(&acc_full[1], 1);
        mbar_init(&acc_empty[0], 128); mbar_init(&acc_empty[1], 128);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(VT_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    auto issue_load = [&](int g) {
        const int ti = g / gpt, grp = g - ti * gpt;
        const int tile = (int)blockIdx.x + ti * (int)gridDim.x;
        const int s = g % NS;
        mbar_expect_tx(&full[s], (uint32_t)kps * VT_KB_BYTES);
        for (int j = 0; j < kps; ++j)
            tma_load_2d(&tmW, &full[s], sA + (size_t)s * VT_STAGE_BYTES + (size_t)j * VT_KB_BYTES, (grp * kps + j) * VT_BK, tile * VT_BM);
    };
    int issued = 0;
    if (warp == 0 && lane == 0) {
        const int first = total_g < NS ? total_g : NS;
        for (; issued < first; ++issued) issue_load(issued);
    }
    const bool has_ln = a.ln_w != nullptr;                                // K <= 512 on this path (checked at launch)
    float4 gw[4], gb[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = i * 128 + lane * 4;
        gw[i] = (has_ln && c < K) ? __ldg(reinterpret_cast<const float4*>(a.ln_w + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
        gb[i] = (has_ln && c < K) ? __ldg(reinterpret_cast<const float4*>(a.ln_b + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    const float* X = a.X;
    if (has_ln) {
        constexpr int NW = VT_THREADS / 32, RPW = (VT_N + NW - 1) / NW;
        float4 xv[RPW][4];
        float s1[RPW], qq[RPW];
#pragma unroll
        for (int j = 0; j < RPW; ++j) {
            const int r = warp + j * NW;
            s1[j] = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int c = i * 128 + lane * 4;
                xv[j][i] = (r < B && c < K) ? __ldcg(reinterpret_cast<const float4*>(X + (size_t)r * K + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int j = 0; j < RPW; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) s1[j] += (xv[j][i].x + xv[j][i].y) + (xv[j][i].z + xv[j][i].w);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int j = 0; j < RPW; ++j) s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], o);
#pragma unroll
        for (int j = 0; j < RPW; ++j) {
            const float mean = s1[j] / (float)K;
            qq[j] = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (i * 128 + lane * 4 < K) {
                    const float t0 = xv[j][i].x - mean, t1 = xv[j][i].y - mean, t2 = xv[j][i].z - mean, t3 = xv[j][i].w - mean;
                    qq[j] += (t0 * t0 + t1 * t1) + (t2 * t2 + t3 * t3);
                }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int j = 0; j < RPW; ++j) qq[j] += __shfl_xor_sync(0xffffffffu, qq[j], o);
#pragma unroll
        for (int j = 0; j < RPW; ++j) {
            const int r = warp + j * NW;
            if (r >= VT_N) continue;
            const float mean = s1[j] / (float)K;
            const float rs = 1.0f / sqrtf(qq[j] / (float)K + 1e-5f);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int c = i * 128 + lane * 4;
                if (c < K) {
                    uint2 pk = make_uint2(0u, 0u);
                    if (r < B) {
                        pk.x = pack_bf16x2((xv[j][i].x - mean) * rs * gw[i].x + gb[i].x, (xv[j][i].y - mean) * rs * gw[i].y + gb[i].y);
                        pk.y = pack_bf16x2((xv[j][i].z - mean) * rs * gw[i].z + gb[i].z, (xv[j][i].w - mean) * rs * gw[i].w + gb[i].w);
                    }
                    const int kb = c >> 6, chunk = (c & 63) >> 3;
                    *reinterpret_cast<uint2*>(sB + (size_t)kb * VT_BTILE_BYTES + r * 128 + ((chunk ^ (r & 7)) << 4) + (c & 7) * 2) = pk;
                }
            }
        }
    } else {
        // plain conversion of X [B][K] f32 -> bf16 swizzled tiles, four 128-bit loads in flight per thread
        const int k4 = K >> 2, total = VT_N * k4;
        for (int base = tid; base < total; base += VT_THREADS * 4) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int idx = base + u * VT_THREADS;
                const int r = idx / k4, c = (idx - r * k4) << 2;
                v[u] = (idx < total && r < B) ? __ldcg(reinterpret_cast<const float4*>(X + (size_t)r * K + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int idx = base + u * VT_THREADS;
                if (idx < total) {
                    const int r = idx / k4, c = (idx - r * k4) << 2;
                    const int kb = c >> 6, chunk = (c & 63) >> 3;
                    *reinterpret_cast<uint2*>(sB + (size_t)kb * VT_BTILE_BYTES + r * 128 + ((chunk ^ (r & 7)) << 4) + (c & 7) * 2) =
                        make_uint2(pack_bf16x2(v[u].x, v[u].y), pack_bf16x2(v[u].z, v[u].w));
                }
            }
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();

    if (warp == 0) {
        if (lane == 0) {
            for (; issued < total_g; ++issued) {
                const int s = issued % NS;
                mbar_wait(&empty[s], (uint32_t)((issued / NS - 1) & 1));
                issue_load(issued);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc();
            int g = 0;
            for (int ti = 0; ti < my_tiles; ++ti) {
                const int acc = ti & 1;
                if (ti >= 2) mbar_wait(&acc_empty[acc], (uint32_t)(((ti >> 1) - 1) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int grp = 0; grp < gpt; ++grp, ++g) {
                    const int s = g % NS;
                    mbar_wait(&full[s], (uint32_t)((g / NS) & 1));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    for (int j = 0; j < kps; ++j) {
                        const int kb = grp * kps + j;
                        const uint64_t da = make_smem_desc(smem_u32(sA + (size_t)s * VT_STAGE_BYTES + (size_t)j * VT_KB_BYTES));
                        const uint64_t db = make_smem_desc(smem_u32(sB + (size_t)kb * VT_BTILE_BYTES));
#pragma unroll
                        for (int k = 0; k < VT_BK / 16; ++k)
                            umma(tmem_base + acc * VT_N, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (uint32_t)((kb | k) != 0));
                    }
                    umma_commit(&empty[s]);
                }
                umma_commit(&acc_full[acc]);
            }
        }
    } else {
        const int q = warp & 3;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        for (int ti = 0; ti < my_tiles; ++ti) {
            const int acc = ti & 1;
            const int n = ((int)blockIdx.x + ti * (int)gridDim.x) * VT_BM + q * 32 + lane;
            const float bias_n = a.bias ? __ldg(a.bias + n) : 0.f;
            float res[VT_N];
            if (a.residual) {
#pragma unroll
                for (int b = 0; b < VT_N; ++b) res[b] = b < B ? __ldcg(a.residual + (size_t)b * N + n) : 0.f;
            }
            mbar_wait(&acc_full[acc], (uint32_t)((ti >> 1) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t r[32];
            tmem_ld32(tmem_base + lane_off + acc * VT_N, r);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(&acc_empty[acc]);
#pragma unroll
            for (int b = 0; b < VT_N; ++b) {
                float v = __uint_as_float(r[b]) + bias_n;
                if (a.act == 1) v = gelu_erf_f(v);
                if (a.residual) v += res[b];
                if (b < B) a.Y[(size_t)b * N + n] = v;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(VT_TMEM_COLS) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

void vocab_tc_alloc(wb_ctx* ctx) {
    const wb_model_cfg& c = ctx->cfg;
    ctx->dec.vocab_tc = nullptr;
    if (c.precision != WB_PREC_BF16 || c.d_model % VT_BK != 0 || c.d_model > VT_MAX_KB * VT_BK || c.max_batch < 1) return;
    if (const char* e = getenv("WB_VOCAB_TC")) if (e[0] == '0') return;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    WB_REQUIRE(p != nullptr && q == cudaDriverEntryPointSuccess, WB_ECUDA, "cuTensorMapEncodeTiled not available");
    auto* v = new VocabTc();
    // tied embedding E [vocab][d] bf16, row-major: tiles of 128 rows x 64 k; rows past the vocabulary arrive as zeros
    cuuint64_t dims[2] = {(cuuint64_t)c.d_model, (cuuint64_t)c.vocab};
    cuuint64_t str[1] = {(cuuint64_t)c.d_model * 2};
    cuuint32_t box[2] = {VT_BK, VT_BM}, estr[2] = {1, 1};
    CUresult r = reinterpret_cast<EncodeTiledFn>(p)(&v->tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ctx->w.embed), dims, str, box, estr,
                                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { delete v; WB_THROW(WB_ECUDA, "cuTensorMapEncodeTiled (vocabulary projection) failed with CUresult %d", (int)r); }
    CUDA_CHECK(cudaFuncSetAttribute(vocab_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)VtCfg<1>::SMEM));
    CUDA_CHECK(cudaFuncSetAttribute(vocab_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)VtCfg<2>::SMEM));
    v->ok = true;
    ctx->dec.vocab_tc = v;
    // the per-layer GEMMs: a tensor map per weight matrix (out % 128 == 0, in % 64 == 0, in <= 2048)
    bool all = true;
    auto add = [&](const LinearW& L) {
        if (!L.w || L.out % VT_BM != 0 || L.in % VT_BK != 0 || L.in > 2048) { all = false; return; }
        CUtensorMap tm;
        cuuint64_t d2[2] = {(cuuint64_t)L.in, (cuuint64_t)L.out};
        cuuint64_t s2[1] = {(cuuint64_t)L.in * 2};
        CUresult rr = reinterpret_cast<EncodeTiledFn>(p)(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, L.w, d2, s2, box, estr,
                                                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rr != CUDA_SUCCESS) { all = false; return; }
        v->linear_maps.emplace_back(L.w, tm);
    };
    for (const auto& L : ctx->w.dec) { add(L.qkv); add(L.o); add(L.cq); add(L.co); add(L.fc1); add(L.fc2); }
    // opt-in (WB_SKINNY_TC=1): measured slower than the mma.sync kernels at batch 32 (4 .. 16 CTAs per GEMM and a TMEM /
    // barrier prologue per launch: decode 43.3 -> 79.3 ms for one batch, 41.2 -> 36.7 k audio-s/s with 8 in flight)
    const char* se = getenv("WB_SKINNY_TC");
    v->skinny_ok = all && se && se[0] == '1';
    if (v->skinny_ok)
        CUDA_CHECK(cudaFuncSetAttribute(skinny_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));   // sized per launch
}

// Enqueues Y = epi(W . f(X)) on the tcgen05 kernel if this weight has a tensor map and the shape fits; false = not taken.
bool skinny_tc_launch(wb_ctx* ctx, cudaStream_t st, bool pdl, const float* X, int B, int K, const void* W, int N, const float* bias,
                      const float* ln_w, const float* ln_b, int act, const float* residual, float* Y) {
    const VocabTc* v = static_cast<const VocabTc*>(ctx->dec.vocab_tc);
    if (!v || !v->skinny_ok || B < 1 || B > VT_N || !Y || N % VT_BM != 0 || K % VT_BK != 0 || K > 2048 || (ln_w && K > 512)) return false;
    const CUtensorMap* tm = nullptr;
    for (const auto& e : v->linear_maps)
        if (e.first == W) { tm = &e.second; break; }
    if (!tm) return false;
    const int nkb = K / VT_BK;
    const size_t fixed = (size_t)nkb * VT_BTILE_BYTES + 1024 + 512;
    int n_stages = (int)((232448 - fixed) / VT_STAGE_BYTES);
    if (n_stages > VT_STAGES) n_stages = VT_STAGES;
    const int gpt = nkb / (nkb % VT_KPS_MAX == 0 ? VT_KPS_MAX : 1);
    if (n_stages > gpt) n_stages = gpt;                                  // one tile per CTA: no more stages than it has loads
    if (n_stages < 1) return false;
    SkinnyTcArgs a{X, B, K, N, ln_w, ln_b, bias, act, residual, Y, n_stages};
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(N / VT_BM); cfg.blockDim = dim3(VT_THREADS);
    cfg.dynamicSmemBytes = (size_t)n_stages * VT_STAGE_BYTES + fixed;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    int na = 0;
    if (pdl) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    CUDA_CHECK(cudaLaunchKernelEx(&cfg, skinny_tc_kernel, *tm, a));
    return true;
}

void vocab_tc_free(wb_ctx* ctx) {
    delete static_cast<VocabTc*>(ctx->dec.vocab_tc);
    ctx->dec.vocab_tc = nullptr;
}

bool vocab_tc_ok(const wb_ctx* ctx, int B) {
    const VocabTc* v = static_cast<const VocabTc*>(ctx->dec.vocab_tc);
    return v && v->ok && B >= 1 && B <= 2 * VT_N;
}

// sequences per partial row written by vocab_tc_launch for a batch of B
int vocab_tc_stride(int B) { return B <= VT_N ? VT_N : 2 * VT_N; }

// Enqueues the kernel; returns the number of per-CTA partials it writes ([cta][vocab_tc_stride(B)] values | indices).
int vocab_tc_launch(wb_ctx* ctx, cudaStream_t st, bool pdl, const int* state, const float* x, int B, float* amax_val, int* amax_idx) {
    const wb_model_cfg& c = ctx->cfg;
    const VocabTc* v = static_cast<const VocabTc*>(ctx->dec.vocab_tc);
    const int n_tiles = (c.vocab + VT_BM - 1) / VT_BM;
    // One CTA per TWO SMs unless the caller said that one batch is in flight (wb_set_load_hint; WB_VOCAB_CTAS overrides): the TMA ring streams a CTA's tiles at a rate that does not
    // need every SM, and with several batches in flight the SM-time a launch holds is what it costs the other batches.
    // Measured (decode only, B = 32; 8 in flight / alone, ms per batch): 148 CTAs 18.78 / 44.5, 74: 18.48 / 44.8,
    // 50: 18.3 / 45.3, 37: 18.2 / 45.7.
    static const int cap = [] { const char* e = getenv("WB_VOCAB_CTAS"); return e ? atoi(e) : 0; }();
    int grid = n_tiles < ctx->sm_count ? n_tiles : ctx->sm_count;
    const int want = cap > 0 ? cap : ctx->dec.load_hint == 1 ? ctx->sm_count : (ctx->sm_count + 1) / 2;   // one batch in flight: every SM
    if (want < grid) grid = want;
    cudaLaunchConfig_t cfg{};
    // WB_VOCAB_L2=none: plain loads (default: evict-last hint, keeps the matrix in L2 under the K/V streams)
    static const int l2_last = [] { const char* e = getenv("WB_VOCAB_L2"); return (e && e[0] == 'n') ? 0 : 1; }();
    const bool two = B > VT_N;
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(VT_THREADS); cfg.dynamicSmemBytes = two ? VtCfg<2>::SMEM : VtCfg<1>::SMEM; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    int na = 0;
    if (pdl) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    CUDA_CHECK(cudaLaunchKernelEx(&cfg, two ? vocab_tc_kernel<2> : vocab_tc_kernel<1>, v->tmW, x, B, c.d_model, c.vocab,
                                  (const float*)ctx->w.dec_ln.w, (const float*)ctx->w.dec_ln.b,
                                  state, (const unsigned*)ctx->dec.sup_base.p, (const unsigned*)ctx->dec.sup_first.p, amax_val, amax_idx, l2_last));
    return grid;
}
