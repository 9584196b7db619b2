// Microbenchmark behind the encoder-attention softmax design (attn_tc.cu): exponentials per clock and SM for
//   ex2.approx.ftz.f32 (MUFU), ex2.approx.ftz.bf16x2 (two per MUFU op?), ex2.approx.f16x2, and a degree-3 polynomial
//   2^x on the FMA pipe (Cody-Waite split, exponent added as an integer).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/ex2 tools/ubench/ex2.cu ; run on the GPU box.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t ex2bf2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t ex2h2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ float ex2poly(float x) {
    x = fmaxf(x, -126.0f);
    const float r = x + 12582912.0f;                  // 1.5 * 2^23: integer part lands in the low mantissa bits
    const float xi = r - 12582912.0f;
    const float f = x - xi;                           // in [-0.5, 0.5]
    float p = fmaf(f, 0.0555041087f, 0.2402265070f);
    p = fmaf(p, f, 0.6931471806f);
    p = fmaf(p, f, 1.0f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(r) << 23));
}

template <int MODE>
__global__ void k(float* out, int iters, float seed) {
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed * (float)(threadIdx.x + i) * 1e-3f - 1.0f;
    uint32_t u[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) u[i] = 0xBF80BF00u + threadIdx.x + i;   // two negative bf16 / f16 values
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) a[i] = ex2f(a[i]) - 1.5f;
            if (MODE == 1) u[i] = ex2bf2(u[i]) ^ 0x80008000u;
            if (MODE == 2) u[i] = ex2h2(u[i]) ^ 0x80008000u;
            if (MODE == 3) a[i] = ex2poly(a[i]) - 1.5f;
            if (MODE == 4) { a[i] = ex2f(a[i]) - 1.5f; a[(i + 4) & 7] = ex2poly(a[(i + 4) & 7]) - 1.5f; }   // both pipes
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(u[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int per_iter_per_thread) {
    float* out; cudaMalloc(&out, 148 * 4 * 1024 * sizeof(float));
    const int iters = 4096;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 4, 256>>>(out, 16, 0.5f);
    cudaEventRecord(e0);
    k<MODE><<<148 * 4, 256>>>(out, iters, 0.5f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double ops = (double)148 * 4 * 256 * iters * per_iter_per_thread;
    printf("%-28s %8.3f ms  %7.1f Gexp/s  %6.2f exp/clk/SM (at %d MHz nominal)\n", name, ms, ops / ms * 1e-6, ops / (ms * 1e-3) / 148 / (clk_khz * 1e3), clk_khz / 1000);
    cudaFree(out);
}

int main() {
    run<0>("ex2.approx.ftz.f32", 8);
    run<1>("ex2.approx.ftz.bf16x2", 16);
    run<2>("ex2.approx.f16x2", 16);
    run<3>("poly3 on FMA pipe", 8);
    run<4>("f32 MUFU + poly3 interleaved", 16);
    // accuracy of the polynomial against exp2f over [-20, 8]
    return 0;
}
