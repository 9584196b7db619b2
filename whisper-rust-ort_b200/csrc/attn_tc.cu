// attn_tc.cu — encoder self-attention of the bf16 build on tcgen05 (kernel K2d): non-causal,
// T = 1500, head_dim 64.  One CTA per (clip, head, 128-query block); per 128-key block:
//   S = Q.K^T        tcgen05.mma 128x128x64 (both operands K-major, TMA 128B swizzle) -> TMEM
//   softmax          8 warps, two threads per query row (64 keys each): tcgen05.ld S, online max/sum in registers,
//                    P -> bf16 -> shared memory in the UMMA K-major swizzled layout
//   O_j = P.V        tcgen05.mma 128x64x128 (V consumed MN-major straight from the TMA tile) -> TMEM
//   O += O_j         in registers (tcgen05.ld), rescaled by exp(m_old - m_new)
// K tiles are double buffered by TMA, V single; two CTAs share an SM (96 KB smem, 256 TMEM columns each),
// so the MMAs of one overlap the softmax of the other, and PV of block j overlaps the softmax of block j+1.
// Replaces the Softmax(QK^T)V sub-graphs ONNX Runtime executes inside encoder.run
// (/root/reference/src/main.rs:703); same math as oracle/whisper_ref.py::_attend.
#include <cuda.h>

#include "ctx.h"

namespace {

constexpr int AQ = 128, AK = 128, HD = 64;
constexpr int ATT_THREADS = 288;                       // warp 0: TMA + MMA issue; warps 1-8: softmax / epilogue
constexpr uint32_t TILE_BYTES = 128 * 64 * 2;          // one [128 x 64] bf16 tile, 128-byte rows
constexpr uint32_t ATT_TMEM_COLS = 256;                // S: [0,128)  O_j: [128,192)
constexpr size_t ATT_SMEM = 6 * TILE_BYTES + 1024 + 128;   // Q, K[2], V, P[2 halves] + barriers: 97 KB -> two CTAs per SM

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
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n"
        "@P1 bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// SWIZZLE_128B descriptor over [rows][64 bf16] tiles (128-byte rows, 8-row groups 1024 B apart).
// Valid both for a K-major operand (rows = M/N index) and for an MN-major operand whose MN extent
// is one 64-element swizzle atom (rows = K index): in both cases the 8-row group stride is SBO.
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)(1024 >> 4) << 16;                  // LBO (only used across MN atoms; single atom here)
    d |= (uint64_t)(1024 >> 4) << 32;                  // SBO
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
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
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// idesc: D f32, A/B bf16, M = 128; N and B-major per use
__device__ __forceinline__ constexpr uint32_t idesc_of(int n, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}


// Softmax halves of one key block for one query row (thread): 64 scores in two tcgen05.ld chunks of 32.
// MASK = the ragged last block (keys >= valid do not exist); full blocks carry no per-score predicate.
template <bool MASK>
__device__ __forceinline__ float row_max64(uint32_t taddr, int half, int valid) {
    float mx = -INFINITY;
#pragma unroll 1
    for (int c2 = 0; c2 < 2; ++c2) {
        const int ch = half * 2 + c2;
        uint32_t r[32];
        tmem_ld32(taddr + ch * 32, r);
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (!MASK || ch * 32 + i < valid) mx = fmaxf(mx, __uint_as_float(r[i]));
    }
    return mx;
}
template <bool MASK>
__device__ __forceinline__ float exp_store64(uint32_t taddr, int half, int valid, float sc, float m_new, uint8_t* prow, int sw) {
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll 1
    for (int c2 = 0; c2 < 2; ++c2) {
        const int ch = half * 2 + c2;
        uint32_t r[32];
        tmem_ld32(taddr + ch * 32, r);
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            float p0 = ex2(fmaf(__uint_as_float(r[2 * i]), sc, -m_new));
            float p1 = ex2(fmaf(__uint_as_float(r[2 * i + 1]), sc, -m_new));
            if (MASK) {
                if (ch * 32 + 2 * i >= valid) p0 = 0.f;
                if (ch * 32 + 2 * i + 1 >= valid) p1 = 0.f;
            }
            sum0 += p0;
            sum1 += p1;
            __nv_bfloat162 pb = __floats2bfloat162_rn(p0, p1);
            pk[i] = *reinterpret_cast<uint32_t*>(&pb);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {                            // 32 keys = chunks c2*4 .. c2*4+3 of this half
            const int chunk = c2 * 4 + c;
            *reinterpret_cast<uint4*>(prow + ((chunk ^ sw) << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
        }
    }
    return sum0 + sum1;
}

__global__ void __launch_bounds__(ATT_THREADS, 2)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, __nv_bfloat16* __restrict__ out, int T, int d, int H, int n_qb) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    uint8_t* sQ = smem;
    uint8_t* sK = smem + TILE_BYTES;                   // 2 buffers
    uint8_t* sV = smem + 3 * TILE_BYTES;               // 1 buffer
    uint8_t* sP = smem + 4 * TILE_BYTES;               // 2 halves of 64 keys
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 6 * TILE_BYTES);
    uint64_t* bar_q = bars;            // 1
    uint64_t* bar_k = bars + 1;        // 2: K tile landed
    uint64_t* bar_kfree = bars + 3;    // 2: S MMA that read the K tile retired
    uint64_t* bar_s = bars + 5;
    uint64_t* bar_p = bars + 6;
    uint64_t* bar_o = bars + 7;        // O_j ready == V tile free
    uint64_t* bar_v = bars + 8;        // V tile landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

    __shared__ float s_xmax[2][2][128];                // [block parity][key half][row]: partial row maxima
    __shared__ float s_xl[2][128];                     // [key half][row]: partial row sums (end of the tile)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qb = blockIdx.x % n_qb, h = (blockIdx.x / n_qb) % H, b = blockIdx.x / (n_qb * H);
    const int n_kb = (T + AK - 1) / AK;

    if (warp == 0) {
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmQKV)) : "memory");
            mbar_init(bar_q, 1);
            mbar_init(&bar_k[0], 1); mbar_init(&bar_k[1], 1);
            mbar_init(&bar_kfree[0], 1); mbar_init(&bar_kfree[1], 1);
            mbar_init(bar_s, 1); mbar_init(bar_p, 256); mbar_init(bar_o, 1); mbar_init(bar_v, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(ATT_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tS = tmem_base, tO = tmem_base + 128;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA + MMA control thread =====
            const uint32_t idesc_s = idesc_of(128, 0), idesc_o = idesc_of(64, 1);
            mbar_expect_tx(bar_q, TILE_BYTES);
            tma_load_3d(&tmQKV, bar_q, sQ, h * HD, qb * AQ, b);
            mbar_expect_tx(&bar_k[0], TILE_BYTES);
            tma_load_3d(&tmQKV, &bar_k[0], sK, d + h * HD, 0, b);
            mbar_expect_tx(bar_v, TILE_BYTES);
            tma_load_3d(&tmQKV, bar_v, sV, 2 * d + h * HD, 0, b);
            mbar_wait(bar_q, 0);
            const uint64_t dq = make_desc_sw128(smem_u32(sQ));
            const uint64_t dv = make_desc_sw128(smem_u32(sV));
            auto issue_s = [&](int j) {                                  // S_j = Q K_j^T into TMEM, then release the K slot
                const int s = j & 1;
                mbar_wait(&bar_k[s], (uint32_t)((j >> 1) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t dk = make_desc_sw128(smem_u32(sK + s * TILE_BYTES));
#pragma unroll
                for (int k = 0; k < HD / 16; ++k)                        // K dim = head_dim
                    umma(tS, dq + (uint64_t)(k * 2), dk + (uint64_t)(k * 2), idesc_s, (uint32_t)(k != 0));
                umma_commit(bar_s);
                umma_commit(&bar_kfree[s]);
            };
            auto load_k = [&](int j) {                                   // K block j into slot j&1 (free once S_{j-2} retired)
                const int s = j & 1;
                if (j >= 2) mbar_wait(&bar_kfree[s], (uint32_t)(((j - 2) >> 1) & 1));
                mbar_expect_tx(&bar_k[s], TILE_BYTES);
                tma_load_3d(&tmQKV, &bar_k[s], sK + s * TILE_BYTES, d + h * HD, j * AK, b);
            };
            issue_s(0);
            if (n_kb > 1) load_k(1);
            for (int j = 0; j < n_kb; ++j) {
                mbar_wait(bar_p, (uint32_t)(j & 1));                     // P_j in smem, S_j and O_{j-1} consumed
                // S_{j+1} goes first: the softmax warps start on it while P_j V_j is still running
                if (j + 1 < n_kb) {
                    issue_s(j + 1);
                    if (j + 2 < n_kb) load_k(j + 2);
                }
                mbar_wait(bar_v, (uint32_t)(j & 1));                     // V_j landed
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int k = 0; k < AK / 16; ++k) {                      // O_j = P_j V_j, K dim = keys
                    const uint64_t dp = make_desc_sw128(smem_u32(sP + (k >> 2) * TILE_BYTES)) + (uint64_t)((k & 3) * 2);
                    umma(tO, dp, dv + (uint64_t)(k * 128), idesc_o, (uint32_t)(k != 0));   // 16 keys = 16 rows x 128 B
                }
                umma_commit(bar_o);
                if (j + 1 < n_kb) {                                      // V tile is free once PV_j retired
                    mbar_wait(bar_o, (uint32_t)(j & 1));
                    mbar_expect_tx(bar_v, TILE_BYTES);
                    tma_load_3d(&tmQKV, bar_v, sV, 2 * d + h * HD, (j + 1) * AK, b);
                }
            }
        }
    } else {
        // ===== softmax + output: 8 warps, two per TMEM lane quadrant; a thread owns one query row and one
        // 64-key half of the block (and 32 of the 64 output columns); the two threads of a row swap their
        // partial row maximum through shared memory + a 64-thread named barrier =====
        const int q = warp & 3;                                          // TMEM lane quadrant of this warp
        const int half = (warp - 1) >> 2;                                // warps 1-4: keys 0-63, warps 5-8: keys 64-127
        const int row = q * 32 + lane;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const float sc = 0.125f * 1.4426950408889634f;                   // head_dim^-0.5 * log2(e)
        float m = -INFINITY, l = 0.f;                                    // l: this thread's half of the row sum
        float o[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = 0.f;
        uint8_t* prow = sP + half * TILE_BYTES + row * 128;
        const int sw = row & 7;

        for (int j = 0; j < n_kb; ++j) {
            mbar_wait(bar_s, (uint32_t)(j & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int valid = T - j * AK;                                // keys of this block that exist
            // pass 1: maximum over this thread's 64 keys, then over the row (only the last key block is ragged)
            const bool full = valid >= AK;
            const float mx_own = full ? row_max64<false>(tS + lane_off, half, valid) : row_max64<true>(tS + lane_off, half, valid);
            float mx = mx_own;
            s_xmax[j & 1][half][row] = mx;
            asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");    // the two warps of this quadrant
            mx = fmaxf(mx, s_xmax[j & 1][half ^ 1][row]);
            const float m_new = fmaxf(m, mx * sc);
            const float alpha = ex2(m - m_new);                          // 0 on the first block (m = -inf)
            // P is single-buffered and S_j was issued BEFORE P_{j-1} V_{j-1}: bar_s(j) does not imply that the tensor
            // core is done reading P_{j-1}.  Wait for that MMA (bar_o) before overwriting P -- normally long complete.
            if (j > 0) mbar_wait(bar_o, (uint32_t)((j - 1) & 1));
            // pass 2: p = exp2(s*sc - m_new) -> bf16 -> swizzled smem; partial row sum in f32
            const float sum = full ? exp_store64<false>(tS + lane_off, half, valid, sc, m_new, prow, sw)
                                   : exp_store64<true>(tS + lane_off, half, valid, sc, m_new, prow, sw);
            l = l * alpha + sum;
            m = m_new;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // P visible to the tensor core
            if (j > 0) {                                                   // fold in O_{j-1} (own 32 columns), then rescale
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                uint32_t r[32];
                tmem_ld32(tO + lane_off + half * 32, r);
#pragma unroll
                for (int i = 0; i < 32; ++i) o[i] = (o[i] + __uint_as_float(r[i])) * alpha;
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(bar_p);
        }
        mbar_wait(bar_o, (uint32_t)((n_kb - 1) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        s_xl[half][row] = l;
        asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
        const float inv = 1.0f / (l + s_xl[half ^ 1][row]);
        const int gq = qb * AQ + row;
        {
            uint32_t r[32];
            tmem_ld32(tO + lane_off + half * 32, r);
            if (gq < T) {
                uint4* dst = reinterpret_cast<uint4*>(out + ((size_t)b * T + gq) * d + h * HD + half * 32);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t w[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int e = c * 8 + 2 * i;
                        __nv_bfloat162 v = __floats2bfloat162_rn((o[e] + __uint_as_float(r[e])) * inv, (o[e + 1] + __uint_as_float(r[e + 1])) * inv);
                        w[i] = *reinterpret_cast<uint32_t*>(&v);
                    }
                    dst[c] = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ATT_TMEM_COLS) : "memory");
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// Second generation of the same kernel (default; WB_ATTN_V=1 selects the first).  Same tiles, barriers and MMA order;
// what changed is the softmax side, which bounded the first one (ncu: ex2 pipe 50 %, issue 52 %, tensor 24 %):
//   * the 64 scores of a thread are read from TMEM ONCE and stay in registers between the maximum and the exponentials
//     (the first version read S twice: tcgen05.ld + wait per 32 columns, 4 of them per block and thread);
//   * O is accumulated by the tensor core in TMEM (accumulate flag on P.V) instead of being read back and folded into
//     32 registers per thread every block;  the running maximum used for the exponentials is only moved when the true
//     maximum has grown by more than 2^8 (exact: l and O carry the same stale maximum and the final O / l cancels it),
//     so the TMEM read-modify-write of O happens a few times per tile instead of every block.
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
          "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
          "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// 2^x on the FMA pipe: round-to-nearest split x = xi + f (|f| <= 0.5) with the 1.5 * 2^23 trick, minimax cubic for 2^f
// (relative error 7.5e-5, below the bf16 rounding of P: 2e-3), xi added straight into the exponent field.  The MUFU unit
// does 16 ex2 per clock and SM, the FMA pipe 128 lanes: moving a fraction of the exponentials here lifts the softmax
// bound of a head-dim-64 attention (one exponential per 128 tensor-core flops).
__device__ __forceinline__ float ex2_poly(float x) {
    x = fmaxf(x, -120.0f);
    const float r = x + 12582912.0f;
    const float f = x - (r - 12582912.0f);
    float p = fmaf(f, 0.05517167f, 0.24261112f);
    p = fmaf(p, f, 0.69326099f);
    p = fmaf(p, f, 0.99992807f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(r) << 23));
}
// PF of every 8 exponentials go to the polynomial, evenly spread
__device__ __forceinline__ constexpr bool poly_slot(int i, int PF) { return ((i + 1) * PF) / 8 != (i * PF) / 8; }

// p = exp2(s * sc - m) of the 64 scores a thread holds in registers -> bf16 pairs in pk[32]; returns their f32 sum.
// (Packing first and storing later lets the caller wait for the P tile to become free AFTER the exponentials.)
template <bool MASK, int PF>
__device__ __forceinline__ float exp_pack_regs(const uint32_t (&r)[64], int valid, float sc, float m, uint32_t (&pk)[32]) {
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const int e = 2 * i;
        const float x0 = fmaf(__uint_as_float(r[e]), sc, -m), x1 = fmaf(__uint_as_float(r[e + 1]), sc, -m);
        float p0 = poly_slot(e & 7, PF) ? ex2_poly(x0) : ex2(x0);
        float p1 = poly_slot((e + 1) & 7, PF) ? ex2_poly(x1) : ex2(x1);
        if (MASK) {
            if (e >= valid) p0 = 0.f;
            if (e + 1 >= valid) p1 = 0.f;
        }
        sum0 += p0;
        sum1 += p1;
        __nv_bfloat162 pb = __floats2bfloat162_rn(p0, p1);
        pk[i] = *reinterpret_cast<uint32_t*>(&pb);
    }
    return sum0 + sum1;
}

__device__ __forceinline__ float max3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

constexpr float ATT_RESCALE_LOG2 = 8.0f;               // the stale maximum may lag the true one by 2^8 (p <= 256: exact in f32 sums, same relative precision in bf16)

template <int PF, bool LATE>
__global__ void __launch_bounds__(ATT_THREADS, 2)
attn_tc2_kernel(const __grid_constant__ CUtensorMap tmQKV, __nv_bfloat16* __restrict__ out, int T, int d, int H, int n_qb) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    uint8_t* sQ = smem;
    uint8_t* sK = smem + TILE_BYTES;                   // 2 buffers
    uint8_t* sV = smem + 3 * TILE_BYTES;               // 1 buffer
    uint8_t* sP = smem + 4 * TILE_BYTES;               // 2 halves of 64 keys
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 6 * TILE_BYTES);
    uint64_t* bar_q = bars;
    uint64_t* bar_k = bars + 1;        // 2
    uint64_t* bar_kfree = bars + 3;    // 2
    uint64_t* bar_s = bars + 5;
    uint64_t* bar_p = bars + 6;
    uint64_t* bar_o = bars + 7;        // P_j V_j retired == V tile and P tile free, O stable
    uint64_t* bar_v = bars + 8;
    uint64_t* bar_sfree = bars + 9;    // all 256 softmax threads hold S_j in registers: the S columns may be overwritten
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

    __shared__ float s_xmax[2][2][128];
    __shared__ float s_xl[2][128];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qb = blockIdx.x % n_qb, h = (blockIdx.x / n_qb) % H, b = blockIdx.x / (n_qb * H);
    const int n_kb = (T + AK - 1) / AK;

    if (warp == 0) {
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmQKV)) : "memory");
            mbar_init(bar_q, 1);
            mbar_init(&bar_k[0], 1); mbar_init(&bar_k[1], 1);
            mbar_init(&bar_kfree[0], 1); mbar_init(&bar_kfree[1], 1);
            mbar_init(bar_s, 1); mbar_init(bar_p, 256); mbar_init(bar_o, 1); mbar_init(bar_v, 1); mbar_init(bar_sfree, 256);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(ATT_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tS = tmem_base, tO = tmem_base + 128;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA + MMA control thread (as in the first version, except that P.V accumulates into O) =====
            const uint32_t idesc_s = idesc_of(128, 0), idesc_o = idesc_of(64, 1);
            mbar_expect_tx(bar_q, TILE_BYTES);
            tma_load_3d(&tmQKV, bar_q, sQ, h * HD, qb * AQ, b);
            mbar_expect_tx(&bar_k[0], TILE_BYTES);
            tma_load_3d(&tmQKV, &bar_k[0], sK, d + h * HD, 0, b);
            mbar_expect_tx(bar_v, TILE_BYTES);
            tma_load_3d(&tmQKV, bar_v, sV, 2 * d + h * HD, 0, b);
            mbar_wait(bar_q, 0);
            const uint64_t dq = make_desc_sw128(smem_u32(sQ));
            const uint64_t dv = make_desc_sw128(smem_u32(sV));
            auto issue_s = [&](int j) {
                const int s = j & 1;
                mbar_wait(&bar_k[s], (uint32_t)((j >> 1) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t dk = make_desc_sw128(smem_u32(sK + s * TILE_BYTES));
#pragma unroll
                for (int k = 0; k < HD / 16; ++k)
                    umma(tS, dq + (uint64_t)(k * 2), dk + (uint64_t)(k * 2), idesc_s, (uint32_t)(k != 0));
                umma_commit(bar_s);
                umma_commit(&bar_kfree[s]);
            };
            auto load_k = [&](int j) {
                const int s = j & 1;
                if (j >= 2) mbar_wait(&bar_kfree[s], (uint32_t)(((j - 2) >> 1) & 1));
                mbar_expect_tx(&bar_k[s], TILE_BYTES);
                tma_load_3d(&tmQKV, &bar_k[s], sK + s * TILE_BYTES, d + h * HD, j * AK, b);
            };
            issue_s(0);
            if (n_kb > 1) load_k(1);
            for (int j = 0; j < n_kb; ++j) {
                // S_{j+1} is issued as soon as S_j sits in the softmax threads' registers, i.e. under the exponentials of
                // block j (ncu on the first cut of this kernel: 20 % of all warp samples were softmax warps waiting for S)
                if (j + 1 < n_kb) {
                    mbar_wait(bar_sfree, (uint32_t)(j & 1));
                    issue_s(j + 1);
                    if (j + 2 < n_kb) load_k(j + 2);
                }
                mbar_wait(bar_p, (uint32_t)(j & 1));                     // P_j in smem, O rescaled if it had to be
                mbar_wait(bar_v, (uint32_t)(j & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int k = 0; k < AK / 16; ++k) {                      // O += P_j V_j
                    const uint64_t dp = make_desc_sw128(smem_u32(sP + (k >> 2) * TILE_BYTES)) + (uint64_t)((k & 3) * 2);
                    umma(tO, dp, dv + (uint64_t)(k * 128), idesc_o, (uint32_t)((j | k) != 0));
                }
                umma_commit(bar_o);
                if (j + 1 < n_kb) {
                    mbar_wait(bar_o, (uint32_t)(j & 1));
                    mbar_expect_tx(bar_v, TILE_BYTES);
                    tma_load_3d(&tmQKV, bar_v, sV, 2 * d + h * HD, (j + 1) * AK, b);
                }
            }
        }
    } else {
        const int q = warp & 3;
        const int half = (warp - 1) >> 2;
        const int row = q * 32 + lane;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const float sc = 0.125f * 1.4426950408889634f;
        float m_used = -INFINITY, l = 0.f;                               // l: this thread's half of the row sum, in units of 2^m_used
        uint8_t* prow = sP + half * TILE_BYTES + row * 128;
        const int sw = row & 7;

        for (int j = 0; j < n_kb; ++j) {
            mbar_wait(bar_s, (uint32_t)(j & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int valid = T - j * AK - half * 64;                    // keys of this thread's half that exist (may be <= 0)
            uint32_t r[64];
            tmem_ld32_nowait(tS + lane_off + half * 64, r);
            tmem_ld32_nowait(tS + lane_off + half * 64 + 32, r + 32);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(bar_sfree);                                      // S_j is in registers
            float mx = -INFINITY;
            if (valid >= 64) {
                float mx2 = -INFINITY;                                   // two chains of 3-input maxima (FMNMX3): 32 instructions for 64 scores
#pragma unroll
                for (int i = 0; i < 64; i += 4) {
                    mx = max3(mx, __uint_as_float(r[i]), __uint_as_float(r[i + 1]));
                    mx2 = max3(mx2, __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
                }
                mx = fmaxf(mx, mx2);
            } else {
#pragma unroll
                for (int i = 0; i < 64; ++i)
                    if (i < valid) mx = fmaxf(mx, __uint_as_float(r[i]));
            }
            s_xmax[j & 1][half][row] = mx;
            asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
            mx = fmaxf(mx, s_xmax[j & 1][half ^ 1][row]);
            const float m_cand = mx * sc;
            // the maximum the exponentials use: moved only when the true one has grown by more than 2^8 (decided here, the
            // matching rescale of O in TMEM waits until P_{j-1} V_{j-1} has retired, below)
            float alpha = 1.0f;
            bool grow = false;
            if (j == 0) {
                m_used = m_cand;
            } else {
                grow = m_cand > m_used + ATT_RESCALE_LOG2;
                if (grow) { alpha = ex2(m_used - m_cand); m_used = m_cand; l *= alpha; }
            }
            uint32_t pk[32];
            float sum = 0.f;
            if (LATE) sum = valid >= 64 ? exp_pack_regs<false, PF>(r, valid, sc, m_used, pk) : exp_pack_regs<true, PF>(r, valid, sc, m_used, pk);
            // P is single-buffered and O must be stable: P_{j-1} V_{j-1} has to retire before either is touched -- by now it
            // has had the 64 exponentials of this block to do so (ncu before this reordering: 9.5 % of the samples here)
            if (j > 0) {
                mbar_wait(bar_o, (uint32_t)((j - 1) & 1));
                if (__any_sync(0xffffffffu, grow)) {                     // rare after the first blocks
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
                    for (int c = 0; c < 4; ++c) {                        // 8 columns at a time: few live registers
                        uint32_t o[8];
                        const uint32_t ta = tO + lane_off + half * 32 + c * 8;
                        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                                     : "=r"(o[0]), "=r"(o[1]), "=r"(o[2]), "=r"(o[3]), "=r"(o[4]), "=r"(o[5]), "=r"(o[6]), "=r"(o[7]) : "r"(ta));
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                        for (int i = 0; i < 8; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                                     ::"r"(ta), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]) : "memory");
                    }
                    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                }
            }
            if (!LATE) sum = valid >= 64 ? exp_pack_regs<false, PF>(r, valid, sc, m_used, pk) : exp_pack_regs<true, PF>(r, valid, sc, m_used, pk);
#pragma unroll
            for (int c = 0; c < 8; ++c)
                *reinterpret_cast<uint4*>(prow + ((c ^ sw) << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
            l += sum;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(bar_p);
        }
        mbar_wait(bar_o, (uint32_t)((n_kb - 1) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        s_xl[half][row] = l;
        asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
        const float inv = 1.0f / (l + s_xl[half ^ 1][row]);
        const int gq = qb * AQ + row;
        {
            uint32_t o[32];
            tmem_ld32(tO + lane_off + half * 32, o);
            if (gq < T) {
                uint4* dst = reinterpret_cast<uint4*>(out + ((size_t)b * T + gq) * d + h * HD + half * 32);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t w[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int e = c * 8 + 2 * i;
                        __nv_bfloat162 v = __floats2bfloat162_rn(__uint_as_float(o[e]) * inv, __uint_as_float(o[e + 1]) * inv);
                        w[i] = *reinterpret_cast<uint32_t*>(&v);
                    }
                    dst[c] = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ATT_TMEM_COLS) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

bool attn_tc_enabled() {
    static int enabled = -1;
    if (enabled < 0) { const char* e = getenv("WB_TC_ATTN"); enabled = !(e && e[0] == '0'); }
    return enabled != 0;
}

void attn_tc_set_attrs() {           // per context / device, from wb_create (see mel_set_attrs)
    CUDA_CHECK(cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ATT_SMEM));
    CUDA_CHECK(cudaFuncSetAttribute(attn_tc2_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ATT_SMEM));
    CUDA_CHECK(cudaFuncSetAttribute(attn_tc2_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ATT_SMEM));
    CUDA_CHECK(cudaFuncSetAttribute(attn_tc2_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ATT_SMEM));
    CUDA_CHECK(cudaFuncSetAttribute(attn_tc2_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ATT_SMEM));
    CUDA_CHECK(cudaFuncSetAttribute(attn_tc2_kernel<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ATT_SMEM));
    CUDA_CHECK(cudaFuncSetAttribute(attn_tc2_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ATT_SMEM));
}

// qkv: [B][T][3d] bf16 (q | k | v), out: [B][T][d] bf16.
void attn_tc(wb_ctx* ctx, const void* qkv, void* out, int B, int T, int d, int H) {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        WB_REQUIRE(p != nullptr && q == cudaDriverEntryPointSuccess, WB_ECUDA, "cuTensorMapEncodeTiled not available");
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    WB_REQUIRE(d == H * HD, WB_EINVAL, "attention kernel needs head_dim 64");
    CUtensorMap tm;
    cuuint64_t dims[3] = {(cuuint64_t)(3 * d), (cuuint64_t)T, (cuuint64_t)B};
    cuuint64_t str[2] = {(cuuint64_t)(3 * d) * 2, (cuuint64_t)T * 3 * d * 2};
    cuuint32_t box[3] = {HD, 128, 1}, estr[3] = {1, 1, 1};
    CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(qkv), dims, str, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    WB_REQUIRE(r == CUDA_SUCCESS, WB_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    const int n_qb = ceil_div(T, AQ);
    // process-wide switches, read once (initialisation of a function-local static is thread-safe)
    struct Sw { int version, poly, late; };
    static const Sw sw = [] {
        Sw r{2, 0, 0};
        const char* e = getenv("WB_ATTN_V");
        if (e && e[0] == '1') r.version = 1;
        const char* pe = getenv("WB_ATTN_POLY");            // exponentials per 8 that go to the FMA-pipe polynomial (0 .. 4)
        r.poly = pe ? atoi(pe) : 0;
        if (r.poly < 0 || r.poly > 4) r.poly = 0;
        const char* le = getenv("WB_ATTN_LATE");            // 1: exponentials before the wait for the P tile
        r.late = (le && le[0] == '1') ? 1 : 0;
        return r;
    }();
    const int version = sw.version, poly = sw.poly, late = sw.late;
    const dim3 grid(B * H * n_qb);
    __nv_bfloat16* o = (__nv_bfloat16*)out;
    if (version == 1) attn_tc_kernel<<<grid, ATT_THREADS, ATT_SMEM, ctx->stream>>>(tm, o, T, d, H, n_qb);
    else if (poly == 1) attn_tc2_kernel<1, false><<<grid, ATT_THREADS, ATT_SMEM, ctx->stream>>>(tm, o, T, d, H, n_qb);
    else if (poly == 2) attn_tc2_kernel<2, false><<<grid, ATT_THREADS, ATT_SMEM, ctx->stream>>>(tm, o, T, d, H, n_qb);
    else if (poly == 3) attn_tc2_kernel<3, false><<<grid, ATT_THREADS, ATT_SMEM, ctx->stream>>>(tm, o, T, d, H, n_qb);
    else if (poly == 4) attn_tc2_kernel<4, false><<<grid, ATT_THREADS, ATT_SMEM, ctx->stream>>>(tm, o, T, d, H, n_qb);
    else if (late) attn_tc2_kernel<0, true><<<grid, ATT_THREADS, ATT_SMEM, ctx->stream>>>(tm, o, T, d, H, n_qb);
    else attn_tc2_kernel<0, false><<<grid, ATT_THREADS, ATT_SMEM, ctx->stream>>>(tm, o, T, d, H, n_qb);
    CUDA_CHECK(cudaGetLastError());
}
