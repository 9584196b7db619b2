// mel_math.h — register-resident FFT building blocks for the log-mel kernel (K1).
// 400 = 20 x 20 (four-step): each of 20 threads runs a 20-point DFT (4 x 5 prime-factor, fully
// unrolled, no inner twiddles), exchanges through shared memory with a W_400 twiddle, and runs a
// second 20-point DFT.  Two real frames are packed as one complex signal (re = frame A,
// im = frame B) and separated by Hermitian symmetry.  Host+device so the math is unit-tested on
// the CPU (wb_selftest_fft400) before it ever meets a GPU.
// Replaces rustfft's 400-point plan (reference call sites /root/reference/src/main.rs:440-441,473).
// Included twice by mel.cu: once as namespace fft_scalar (one FADD per component) and once, with WB_PACKED_F32X2
// defined, as fft_packed (sm_100 FADD2: both components of a complex add in one instruction).
#ifndef WB_MEL_MATH_COMMON
#define WB_MEL_MATH_COMMON
#ifdef __CUDA_ARCH__
#define WB_UNROLL _Pragma("unroll")
#else
#define WB_UNROLL            /* host build of the same code (wb_selftest_fft400) */
#endif
#ifdef __CUDACC__
#define WB_HD __host__ __device__ __forceinline__
#else
#define WB_HD inline
#endif

struct c32 {
    float x, y;
};
#endif

#ifdef WB_PACKED_F32X2
namespace fft_packed {
#else
namespace fft_scalar {
#endif
#if defined(__CUDA_ARCH__) && defined(WB_PACKED_F32X2)
// sm_100: one FADD2 adds both halves of a complex number
WB_HD c32 cadd(c32 a, c32 b) { float2 r = __fadd2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y)); return {r.x, r.y}; }
WB_HD c32 csub(c32 a, c32 b) { float2 r = __fadd2_rn(make_float2(a.x, a.y), make_float2(-b.x, -b.y)); return {r.x, r.y}; }
#else
WB_HD c32 cadd(c32 a, c32 b) { return {a.x + b.x, a.y + b.y}; }
WB_HD c32 csub(c32 a, c32 b) { return {a.x - b.x, a.y - b.y}; }
#endif
WB_HD c32 cmul(c32 a, c32 b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
WB_HD c32 mul_negi(c32 a) { return {a.y, -a.x}; }   // (-i) * a
WB_HD c32 mul_posi(c32 a) { return {-a.y, a.x}; }   // (+i) * a

// Forward radix-4: X[c] = sum_a x[a] * (-i)^(a*c)
WB_HD void dft4(c32& x0, c32& x1, c32& x2, c32& x3) {
    c32 s02 = cadd(x0, x2), d02 = csub(x0, x2);
    c32 s13 = cadd(x1, x3), d13 = csub(x1, x3);
    c32 nd = mul_negi(d13);
    x0 = cadd(s02, s13);
    x1 = cadd(d02, nd);
    x2 = csub(s02, s13);
    x3 = csub(d02, nd);
}

// Forward radix-5: X[e] = sum_b x[b] * exp(-2*pi*i*b*e/5)
WB_HD void dft5(c32& x0, c32& x1, c32& x2, c32& x3, c32& x4) {
    const float c1 = 0.30901699437494742f, c2 = -0.80901699437494742f;
    const float s1 = 0.95105651629515357f, s2 = 0.58778525229247313f;
    c32 t1 = cadd(x1, x4), t2 = cadd(x2, x3), t3 = csub(x1, x4), t4 = csub(x2, x3);
    c32 m1 = {x0.x + c1 * t1.x + c2 * t2.x, x0.y + c1 * t1.y + c2 * t2.y};
    c32 m2 = {x0.x + c2 * t1.x + c1 * t2.x, x0.y + c2 * t1.y + c1 * t2.y};
    c32 n1 = {s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y};
    c32 n2 = {s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y};
    c32 in1 = mul_negi(n1), in2 = mul_negi(n2);
    x0 = cadd(x0, cadd(t1, t2));
    x1 = cadd(m1, in1);
    x4 = csub(m1, in1);
    x2 = cadd(m2, in2);
    x3 = csub(m2, in2);
}

// In-place forward 20-point DFT, prime-factor (Good-Thomas) form: 4 and 5 are coprime, so with
//   n = (5 n1 + 4 n2) mod 20   and   k = (5 k1 + 16 k2) mod 20
// W_20^(nk) = W_4^(n1 k1) * W_5^(n2 k2): five radix-4 and four radix-5 butterflies and NO twiddles in between.
// Every index is a compile-time constant after unrolling, so v[] stays in registers.
WB_HD void dft20(c32 (&v)[20]) {
    c32 u[4][5];
WB_UNROLL
    for (int n2 = 0; n2 < 5; ++n2) {
        c32 a0 = v[(4 * n2) % 20], a1 = v[(5 + 4 * n2) % 20], a2 = v[(10 + 4 * n2) % 20], a3 = v[(15 + 4 * n2) % 20];
        dft4(a0, a1, a2, a3);
        u[0][n2] = a0;
        u[1][n2] = a1;
        u[2][n2] = a2;
        u[3][n2] = a3;
    }
WB_UNROLL
    for (int k1 = 0; k1 < 4; ++k1) {
        c32 y0 = u[k1][0], y1 = u[k1][1], y2 = u[k1][2], y3 = u[k1][3], y4 = u[k1][4];
        dft5(y0, y1, y2, y3, y4);
        v[(5 * k1) % 20] = y0;
        v[(5 * k1 + 16) % 20] = y1;
        v[(5 * k1 + 32) % 20] = y2;
        v[(5 * k1 + 48) % 20] = y3;
        v[(5 * k1 + 64) % 20] = y4;
    }
}
}  // namespace fft_scalar / fft_packed
