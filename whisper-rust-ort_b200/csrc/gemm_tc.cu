// gemm_tc.cu — the bf16 tensor-core GEMM of the fast build: C = epi(A[M,K] . W[N,K]^T), fp32
// accumulation in TMEM.  Hand-written for sm_100a: TMA (cp.async.bulk.tensor, 128-byte swizzle)
// feeds a 5-stage shared-memory ring, one elected thread issues tcgen05.mma (cta_group::1,
// kind::f16, 128x128x16), accumulators live in TMEM (two 128-column buffers so the epilogue of
// tile i overlaps the main loop of tile i+1), four epilogue warps drain them with tcgen05.ld and
// apply bias / exact GELU / positional add / f32 residual before the store.  Persistent: one CTA
// per SM walks output tiles (n fastest, so concurrently running CTAs share A rows in L2).
//
// Serves every weight GEMM of the encoder and the cross-attention K/V projection, i.e. what ONNX
// Runtime's MatMul/Gemm/Conv kernels compute inside encoder.run / decoder.run
// (/root/reference/src/main.rs:703, 773).  The conv stem reaches it through a 3-D tensor map whose
// rows overlap in memory (row t = 3 consecutive frames), so there is no im2col buffer.
#include <cuda.h>

#include "ctx.h"

namespace {

// Output tile 128 x BN with BN = 128 or 256.  The encoder GEMMs run at the L2 -> SM throughput cap with 128x128
// tiles (ncu: 10.7-13.7 TB/s of TMA reads, profiles/r1_gemm_tc_v5_raw.csv), so the wider tile is the lever: per
// k-block it loads 16 KB of A + 32 KB of B for twice the flops (87 instead of 64 flop per byte from L2).
constexpr int BM = 128, BK = 64, UMMA_K = 16;
constexpr int TC_THREADS = 320;                              // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (2 per TMEM lane quadrant)
constexpr uint32_t STAGE_A = BM * BK * 2;
constexpr uint32_t STAGE_C = BM * 64 * 2;                    // one 128 x 64 bf16 sub-tile (TMA-store staging)
template <int BN> struct TileCfg {
    static constexpr int STAGES = BN == 128 ? 5 : 3;         // 160 KB / 144 KB of operands in flight
    static constexpr uint32_t STAGE_B = BN * BK * 2;
    static constexpr uint32_t TMEM_COLS = 2 * BN;            // 2 accumulators x BN fp32 columns
    static constexpr int SUB = BN / 128;                     // 64-column staging sub-tiles per column half
    static constexpr size_t SMEM = (size_t)STAGES * (STAGE_A + STAGE_B) + 2 * SUB * STAGE_C + 1024 /*align*/ + 256 /*barriers*/;
};

struct TcArgs {
    void* C;
    int tc;                     // WB_F32 | WB_BF16
    int M, N, K, ldc;
    long long sC;               // per-batch stride of C (elements)
    const float* bias;
    int act;
    const float* rowadd;
    int ld_rowadd;
    const float* residual;
    int tiles_m, tiles_n, num_tiles, num_kb;
    int tma_store;              // bf16 output without residual: tiles leave through shared memory + TMA (coalesced)
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
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor: rows of 128 bytes, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);            // start address
    d |= (uint64_t)1 << 16;                            // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset: 8 rows x 128 B
    d |= (uint64_t)1 << 46;                            // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                            // SWIZZLE_128B
    return d;
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N=BN
template <int BN>
__device__ __forceinline__ constexpr uint32_t make_idesc() {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
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
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
// GELU for bf16 outputs: the tanh form on the hardware tanh (one MUFU + 5 FP32 ops per element instead of the
// rational erf's two MUFU + 14).  |tanh-form - erf-form| <= 5e-4 and tanh.approx adds ~2^-11 relative: both are
// below the bf16 rounding of the stored activation; the fp32 validation build keeps the exact erf.
__device__ __forceinline__ float gelu_fast(float x) {
    const float u = x * fmaf(0.0356774081f, x * x, 0.7978845608f);       // sqrt(2/pi) * (x + 0.044715 x^3)
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
    const float hx = 0.5f * x;
    return fmaf(hx, t, hx);
}

template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const TcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);          // SWIZZLE_128B wants 1024-byte alignment
    constexpr int STAGES = TileCfg<BN>::STAGES, SUB = TileCfg<BN>::SUB, NCH = BN / 64;   // NCH: 32-column chunks per thread
    constexpr uint32_t STAGE_B = TileCfg<BN>::STAGE_B, TMEM_COLS = TileCfg<BN>::TMEM_COLS;
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * STAGE_A;
    uint8_t* sC = smem + STAGES * (STAGE_A + STAGE_B);                    // 2*SUB x [128 rows][128 B], 128B-swizzled
    uint64_t* full = reinterpret_cast<uint64_t*>(sC + 2 * SUB * STAGE_C);
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = empty + STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 256); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < a.num_tiles; t += gridDim.x) {
                const int nb = t % a.tiles_n, rest = t / a.tiles_n, mb = rest % a.tiles_m, z = rest / a.tiles_m;
                for (int kb = 0; kb < a.num_kb; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_expect_tx(&full[stage], STAGE_A + STAGE_B);
                    tma_load_3d(&tmA, &full[stage], sA + stage * STAGE_A, kb * BK, mb * BM, z);
                    tma_load_2d(&tmB, &full[stage], sB + stage * STAGE_B, kb * BK, nb * BN);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            const uint32_t idesc = make_idesc<BN>();
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int t = blockIdx.x; t < a.num_tiles; t += gridDim.x, ++it) {
                const int acc = it & 1;
                const uint32_t use = (uint32_t)(it >> 1);
                mbar_wait(&tempty[acc], (use & 1) ^ 1);                  // epilogue drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                for (int kb = 0; kb < a.num_kb; ++kb) {
                    mbar_wait(&full[stage], phase);                      // TMA bytes landed
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint64_t da = make_smem_desc(smem_u32(sA + stage * STAGE_A));
                    const uint64_t db = make_smem_desc(smem_u32(sB + stage * STAGE_B));
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k)               // +32 bytes per K step inside the swizzle atom
                        umma(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (uint32_t)((kb | k) != 0));
                    umma_commit(&empty[stage]);                          // frees the smem slot when the MMAs retire
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tfull[acc]);                                // accumulator ready for the epilogue
            }
        }
    } else {
        // ===== epilogue: TMEM -> registers -> bias/GELU/pos/residual -> global =====
        const int q = warp & 3;                                          // TMEM lane quadrant this warp may access
        const int chalf = (warp - 2) >> 2;                               // which 64-column half of the tile
        const int row_in_tile = q * 32 + lane;
        int it = 0;
        for (int t = blockIdx.x; t < a.num_tiles; t += gridDim.x, ++it) {
            const int nb = t % a.tiles_n, rest = t / a.tiles_n, mb = rest % a.tiles_m, z = rest / a.tiles_m;
            const int acc = it & 1;
            const uint32_t use = (uint32_t)(it >> 1);
            mbar_wait(&tfull[acc], use & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int gm = mb * BM + row_in_tile;
            const bool row_ok = gm < a.M;
            const long long crow = (long long)z * a.sC + (long long)gm * a.ldc;
            if (a.tma_store) {
                // bf16 tile out through shared memory: a thread owns one row x 64 columns = eight 16-byte chunks, written
                // in the 128B-swizzle pattern (conflict-free), then one TMA store per 128x64 half; rows >= M are clipped
                // by the tensor map.  Replaces 32 scattered 16-byte global stores per warp instruction.
                const bool issuer = ((warp - 2) & 3) == 0 && lane == 0;   // one thread per column half
                uint8_t* sCh = sC + chalf * SUB * STAGE_C;
                if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // previous tile has left the staging
                asm volatile("bar.sync %0, 128;" ::"r"(1 + chalf) : "memory");
#pragma unroll 1
                for (int sub = 0; sub < SUB; ++sub) {
                    uint32_t pk[32];
#pragma unroll
                    for (int c2 = 0; c2 < 2; ++c2) {
                        const int ch = chalf * NCH + sub * 2 + c2;
                        uint32_t r[32];
                        tmem_ld32(tmem_base + (uint32_t)(acc * BN + ch * 32) + ((uint32_t)(q * 32) << 16), r);
                        const int gn0 = nb * BN + ch * 32;
                        float v[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                        if (a.bias) {
                            const float4* bp = reinterpret_cast<const float4*>(a.bias + gn0);
#pragma unroll
                            for (int j = 0; j < 8; ++j) { float4 t4 = __ldg(bp + j); v[4 * j] += t4.x; v[4 * j + 1] += t4.y; v[4 * j + 2] += t4.z; v[4 * j + 3] += t4.w; }
                        }
                        if (a.act == 1) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = gelu_fast(v[j]);
                        }
                        if (a.rowadd && row_ok) {
                            const float4* p = reinterpret_cast<const float4*>(a.rowadd + (long long)gm * a.ld_rowadd + gn0);
#pragma unroll
                            for (int j = 0; j < 8; ++j) { float4 t4 = __ldg(p + j); v[4 * j] += t4.x; v[4 * j + 1] += t4.y; v[4 * j + 2] += t4.z; v[4 * j + 3] += t4.w; }
                        }
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            __nv_bfloat162 pb = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
                            pk[c2 * 16 + j] = *reinterpret_cast<unsigned*>(&pb);
                        }
                    }
                    uint8_t* prow = sCh + sub * STAGE_C + row_in_tile * 128;
                    const int sw = row_in_tile & 7;
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        *reinterpret_cast<uint4*>(prow + ((c ^ sw) << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                mbar_arrive(&tempty[acc]);                                // accumulator drained
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("bar.sync %0, 128;" ::"r"(1 + chalf) : "memory");
                if (issuer) {
#pragma unroll
                    for (int sub = 0; sub < SUB; ++sub)
                        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                                     ::"l"(reinterpret_cast<uint64_t>(&tmC)), "r"(smem_u32(sCh + sub * STAGE_C)),
                                       "r"(nb * BN + chalf * (BN / 2) + sub * 64), "r"(mb * BM), "r"(z)
                                     : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                continue;
            }
#pragma unroll 1
            for (int ch = chalf * NCH; ch < (chalf + 1) * NCH; ++ch) {
                uint32_t r[32];
                tmem_ld32(tmem_base + (uint32_t)(acc * BN + ch * 32) + ((uint32_t)(q * 32) << 16), r);
                const int gn0 = nb * BN + ch * 32;
                if (row_ok) {
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                    if (a.bias) {
                        const float4* bp = reinterpret_cast<const float4*>(a.bias + gn0);
#pragma unroll
                        for (int j = 0; j < 8; ++j) { float4 t4 = __ldg(bp + j); v[4 * j] += t4.x; v[4 * j + 1] += t4.y; v[4 * j + 2] += t4.z; v[4 * j + 3] += t4.w; }
                    }
                    if (a.act == 1) {
                        // bf16 outputs: tanh-form GELU on tanh.approx (gelu_fast above: within 5e-4 of the reference's erf
                        // form, below the bf16 rounding of the stored value); f32 outputs keep the exact erff
                        if (a.tc == WB_BF16) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = gelu_fast(v[j]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
                        }
                    }
                    if (a.rowadd) {
                        const float4* p = reinterpret_cast<const float4*>(a.rowadd + (long long)gm * a.ld_rowadd + gn0);
#pragma unroll
                        for (int j = 0; j < 8; ++j) { float4 t4 = __ldg(p + j); v[4 * j] += t4.x; v[4 * j + 1] += t4.y; v[4 * j + 2] += t4.z; v[4 * j + 3] += t4.w; }
                    }
                    if (a.residual) {
                        const float4* p = reinterpret_cast<const float4*>(a.residual + crow + gn0);
#pragma unroll
                        for (int j = 0; j < 8; ++j) { float4 t4 = p[j]; v[4 * j] += t4.x; v[4 * j + 1] += t4.y; v[4 * j + 2] += t4.z; v[4 * j + 3] += t4.w; }
                    }
                    if (a.tc == WB_F32) {
                        float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(a.C) + crow + gn0);
#pragma unroll
                        for (int j = 0; j < 8; ++j) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    } else {
                        uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.C) + crow + gn0);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            __nv_bfloat162 p0 = __floats2bfloat162_rn(v[8 * j], v[8 * j + 1]), p1 = __floats2bfloat162_rn(v[8 * j + 2], v[8 * j + 3]);
                            __nv_bfloat162 p2 = __floats2bfloat162_rn(v[8 * j + 4], v[8 * j + 5]), p3 = __floats2bfloat162_rn(v[8 * j + 6], v[8 * j + 7]);
                            uint4 u;
                            u.x = *reinterpret_cast<unsigned*>(&p0); u.y = *reinterpret_cast<unsigned*>(&p1);
                            u.z = *reinterpret_cast<unsigned*>(&p2); u.w = *reinterpret_cast<unsigned*>(&p3);
                            o[j] = u;
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(&tempty[acc]);                                    // 256 arrivals release the accumulator
        }
        if (a.tma_store) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // staging must outlive its last store
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        WB_REQUIRE(p != nullptr && q == cudaDriverEntryPointSuccess, WB_ECUDA, "cuTensorMapEncodeTiled not available from the driver");
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

void make_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes, const cuuint32_t* box) {
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    WB_REQUIRE(r == CUDA_SUCCESS, WB_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
}

}  // namespace

bool tma_store_enabled() {                          // WB_TC_TMA_STORE=0: direct global stores from the epilogue warps
    static int enabled = -1;
    if (enabled < 0) { const char* e = getenv("WB_TC_TMA_STORE"); enabled = !(e && e[0] == '0'); }
    return enabled != 0;
}

bool gemm_tc_eligible(const GemmArgs& g) {
    static int enabled = -1;
    if (enabled < 0) { const char* e = getenv("WB_TC"); enabled = !(e && e[0] == '0'); }
    return enabled && g.ta == WB_BF16 && g.tb == WB_BF16 && !g.b_kn && g.alpha == 1.0f && g.N % 128 == 0 && g.K % 8 == 0 &&
           g.lda % 8 == 0 && g.ldb % 8 == 0 && g.ldc % 8 == 0 && g.sAi == 0 && g.sBo == 0 && g.sBi == 0 && g.sCi == 0 &&
           (g.batch == 1 || (g.sAo % 8 == 0 && g.inner <= 1)) && (g.ld_rowadd % 4 == 0);
}

void gemm_tc_set_attrs() {           // per context / device, from wb_create (see mel_set_attrs)
    CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TileCfg<128>::SMEM));
    CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TileCfg<256>::SMEM));
}

void gemm_tc(wb_ctx* ctx, const GemmArgs& g) {
    static int wide = -1;                           // WB_TC_BN256=0: 128-wide tiles everywhere
    if (wide < 0) { const char* e = getenv("WB_TC_BN256"); wide = !(e && e[0] == '0'); }
    // wide tiles where there are enough of them: at N = 512 the 128x256 grid is 5.07 waves of 148 CTAs and the
    // out-projection got slower under ncu (81 -> 89 us), fc2 unchanged; N >= 1024 gained 25-40 % (qkv 111 -> 77 us)
    const int BN = (wide && g.N % 256 == 0 && g.N >= 1024) ? 256 : 128;
    CUtensorMap tmA, tmB, tmC;
    const bool tma_store = g.tc == WB_BF16 && g.residual == nullptr && tma_store_enabled();
    {
        cuuint64_t dims[3] = {(cuuint64_t)g.K, (cuuint64_t)g.M, (cuuint64_t)g.batch};
        cuuint64_t str[2] = {(cuuint64_t)g.lda * 2, (cuuint64_t)(g.batch > 1 ? g.sAo : (long long)g.M * g.lda) * 2};
        cuuint32_t box[3] = {BK, BM, 1};
        make_map(&tmA, g.A, 3, dims, str, box);
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)g.K, (cuuint64_t)g.N};
        cuuint64_t str[1] = {(cuuint64_t)g.ldb * 2};
        cuuint32_t box[2] = {BK, (cuuint32_t)BN};
        make_map(&tmB, g.B, 2, dims, str, box);
    }
    if (tma_store) {
        cuuint64_t dims[3] = {(cuuint64_t)g.N, (cuuint64_t)g.M, (cuuint64_t)g.batch};
        cuuint64_t str[2] = {(cuuint64_t)g.ldc * 2, (cuuint64_t)(g.batch > 1 ? g.sCo : (long long)g.M * g.ldc) * 2};
        cuuint32_t box[3] = {64, BM, 1};
        make_map(&tmC, g.C, 3, dims, str, box);
    } else {
        tmC = tmA;                                  // unused
    }
    TcArgs a{};
    a.tma_store = tma_store ? 1 : 0;
    a.C = g.C; a.tc = g.tc; a.M = g.M; a.N = g.N; a.K = g.K; a.ldc = g.ldc; a.sC = g.sCo;
    a.bias = g.bias; a.act = g.act; a.rowadd = g.rowadd; a.ld_rowadd = g.ld_rowadd; a.residual = g.residual;
    a.tiles_m = ceil_div(g.M, BM); a.tiles_n = g.N / BN; a.num_tiles = a.tiles_m * a.tiles_n * g.batch;
    a.num_kb = ceil_div(g.K, BK);
    const int grid = a.num_tiles < ctx->sm_count ? a.num_tiles : ctx->sm_count;
    if (BN == 256) gemm_tc_kernel<256><<<grid, TC_THREADS, TileCfg<256>::SMEM, ctx->stream>>>(tmA, tmB, tmC, a);
    else gemm_tc_kernel<128><<<grid, TC_THREADS, TileCfg<128>::SMEM, ctx->stream>>>(tmA, tmB, tmC, a);
    CUDA_CHECK(cudaGetLastError());
}
