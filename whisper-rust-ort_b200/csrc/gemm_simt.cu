// gemm_simt.cu — fp32-accumulate SIMT GEMM with fused epilogues.  This is the arithmetic of the
// fp32 VALIDATION build (token-identical to the CPU oracle) and the bring-up reference for the
// tcgen05 path in gemm_tc.cu; it stands where ONNX Runtime's MLAS sgemm stands in the reference
// (every MatMul/Gemm/Conv node executed by encoder.run / decoder.run, main.rs:703, 773, 814).
//
// C[M,N] = epi(alpha * A[M,K] . B^T), A rows K-contiguous (lda), B either [N][K] (weights as HF
// stores them) or [K][N] (b_kn: the V operand of P.V).  128x128x16 tiles, 256 threads, 8x8
// register blocking, register-prefetch double buffering.  Operands may be f32 or bf16 (converted
// on the way into shared memory); batched over blockIdx.z with (outer, inner) strides so
// per-(clip, head) attention GEMMs and per-clip conv-stem GEMMs are single launches.
#include "ctx.h"

namespace {

constexpr int BM = 128, BN = 128, BK = 16, NT = 256, LDS = BM + 4;

template <typename T> struct Vec;
template <> struct Vec<float> {
    static constexpr int N = 4;
    static __device__ __forceinline__ void load(const float* p, bool ok, float (&f)[4]) {
        float4 v = ok ? *reinterpret_cast<const float4*>(p) : make_float4(0.f, 0.f, 0.f, 0.f);
        f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
    }
};
template <> struct Vec<__nv_bfloat16> {
    static constexpr int N = 8;
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, bool ok, float (&f)[8]) {
        uint4 v = ok ? *reinterpret_cast<const uint4*>(p) : make_uint4(0u, 0u, 0u, 0u);
        const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            f[2 * i] = __uint_as_float(w[i] << 16);
            f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

__device__ __forceinline__ void store4(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void store4(__nv_bfloat16* p, const float (&v)[4]) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 u;
    u.x = *reinterpret_cast<unsigned*>(&a);
    u.y = *reinterpret_cast<unsigned*>(&b);
    *reinterpret_cast<uint2*>(p) = u;
}

struct DevArgs {
    const void* A; const void* B; void* C;
    int M, N, K, lda, ldb, ldc, inner;
    long long sAo, sAi, sBo, sBi, sCo, sCi;
    float alpha;
    const float* bias; int act;
    const float* rowadd; int ld_rowadd;
    const float* residual;
};

template <typename TA, typename TB, typename TC, bool B_KN>
__global__ void __launch_bounds__(NT) gemm_simt_kernel(DevArgs a) {
    __shared__ __align__(16) float As[2][BK][LDS];
    __shared__ __align__(16) float Bs[2][BK][LDS];
    const int tid = threadIdx.x;
    const int z = blockIdx.z, zo = z / a.inner, zi = z - zo * a.inner;
    const TA* A = reinterpret_cast<const TA*>(a.A) + zo * a.sAo + zi * a.sAi;
    const TB* B = reinterpret_cast<const TB*>(a.B) + zo * a.sBo + zi * a.sBi;
    const long long coff = zo * a.sCo + zi * a.sCi;
    TC* C = reinterpret_cast<TC*>(a.C) + coff;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

    constexpr int VA = Vec<TA>::N, VB = Vec<TB>::N;
    constexpr int NVA = BM * BK / VA / NT;                  // vectors per thread (A)
    constexpr int NVB = BN * BK / VB / NT;
    float ra[NVA][VA], rb[NVB][VB];

    auto load_tiles = [&](int k0) {
#pragma unroll
        for (int v = 0; v < NVA; ++v) {
            int idx = tid + v * NT, row = idx / (BK / VA), kv = idx % (BK / VA);
            int gm = m0 + row, gk = k0 + kv * VA;
            bool ok = gm < a.M && gk < a.K;
            Vec<TA>::load(A + (long long)gm * a.lda + gk, ok, ra[v]);
        }
#pragma unroll
        for (int v = 0; v < NVB; ++v) {
            int idx = tid + v * NT;
            if (!B_KN) {
                int row = idx / (BK / VB), kv = idx % (BK / VB);
                int gn = n0 + row, gk = k0 + kv * VB;
                bool ok = gn < a.N && gk < a.K;
                Vec<TB>::load(B + (long long)gn * a.ldb + gk, ok, rb[v]);
            } else {
                int kr = idx / (BN / VB), nv = idx % (BN / VB);
                int gk = k0 + kr, gn = n0 + nv * VB;
                bool ok = gk < a.K && gn < a.N;
                Vec<TB>::load(B + (long long)gk * a.ldb + gn, ok, rb[v]);
            }
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int v = 0; v < NVA; ++v) {
            int idx = tid + v * NT, row = idx / (BK / VA), kv = idx % (BK / VA);
#pragma unroll
            for (int j = 0; j < VA; ++j) As[buf][kv * VA + j][row] = ra[v][j];
        }
#pragma unroll
        for (int v = 0; v < NVB; ++v) {
            int idx = tid + v * NT;
            if (!B_KN) {
                int row = idx / (BK / VB), kv = idx % (BK / VB);
#pragma unroll
                for (int j = 0; j < VB; ++j) Bs[buf][kv * VB + j][row] = rb[v][j];
            } else {
                int kr = idx / (BN / VB), nv = idx % (BN / VB);
#pragma unroll
                for (int j = 0; j < VB; ++j) Bs[buf][kr][nv * VB + j] = rb[v][j];
            }
        }
    };

    const int tx = tid % 16, ty = tid / 16;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

    const int nk = (a.K + BK - 1) / BK;
    load_tiles(0);
    store_tiles(0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) load_tiles((kt + 1) * BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
            float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
            float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
            float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
            store_tiles(buf ^ 1);
            __syncthreads();
        }
    }

    // ---- epilogue: alpha, bias, GELU, positional add, residual ----
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (gm >= a.M) continue;
#pragma unroll
        for (int jh = 0; jh < 2; ++jh) {
            const int gn = n0 + jh * 64 + tx * 4;
            if (gn >= a.N) continue;
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float t = a.alpha * acc[i][jh * 4 + j];
                if (a.bias) t += a.bias[gn + j];
                if (a.act == 1) t = gelu_erf(t);
                if (a.rowadd) t += a.rowadd[(long long)gm * a.ld_rowadd + gn + j];
                if (a.residual) t += a.residual[coff + (long long)gm * a.ldc + gn + j];
                v[j] = t;
            }
            store4(C + (long long)gm * a.ldc + gn, v);
        }
    }
}

template <typename TA, typename TB, typename TC>
void launch(wb_ctx* ctx, const GemmArgs& g, const DevArgs& d) {
    dim3 grid(ceil_div(g.N, BN), ceil_div(g.M, BM), g.batch);
    if (g.b_kn) gemm_simt_kernel<TA, TB, TC, true><<<grid, NT, 0, ctx->stream>>>(d);
    else gemm_simt_kernel<TA, TB, TC, false><<<grid, NT, 0, ctx->stream>>>(d);
    CUDA_CHECK(cudaGetLastError());
}

}  // namespace

void gemm_simt(wb_ctx* ctx, const GemmArgs& g) {
    const int va = g.ta == WB_BF16 ? 8 : 4, vb = g.tb == WB_BF16 ? 8 : 4;
    WB_REQUIRE(g.K % va == 0 && g.lda % va == 0, WB_EINVAL, "gemm: A alignment (K=%d lda=%d)", g.K, g.lda);
    if (g.b_kn) WB_REQUIRE(g.N % vb == 0 && g.ldb % vb == 0, WB_EINVAL, "gemm: B[K][N] alignment (N=%d ldb=%d)", g.N, g.ldb);
    else WB_REQUIRE(g.K % vb == 0 && g.ldb % vb == 0, WB_EINVAL, "gemm: B alignment (K=%d ldb=%d)", g.K, g.ldb);
    WB_REQUIRE(g.N % 4 == 0 && g.ldc % 4 == 0, WB_EINVAL, "gemm: C alignment (N=%d ldc=%d)", g.N, g.ldc);
    DevArgs d{g.A, g.B, g.C, g.M, g.N, g.K, g.lda, g.ldb, g.ldc, g.inner < 1 ? 1 : g.inner,
              g.sAo, g.sAi, g.sBo, g.sBi, g.sCo, g.sCi, g.alpha, g.bias, g.act, g.rowadd, g.ld_rowadd, g.residual};
    const int key = g.ta * 100 + g.tb * 10 + g.tc;
    using bf = __nv_bfloat16;
    switch (key) {
        case 0: launch<float, float, float>(ctx, g, d); break;
        case 111: launch<bf, bf, bf>(ctx, g, d); break;
        case 110: launch<bf, bf, float>(ctx, g, d); break;
        case 11: launch<float, bf, bf>(ctx, g, d); break;
        default: WB_THROW(WB_EINVAL, "gemm_simt: unsupported dtype combination %d", key);
    }
}
