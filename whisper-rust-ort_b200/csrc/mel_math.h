// mel_math.h — register-resident FFT building blocks for the log-mel kernel (K1).
// 400 = 20 x 20 (four-step): each of 20 threads runs a 20-point DFT (4 x 5 Cooley-Tukey, fully
// unrolled, constants folded), exchanges through shared memory with a W_400 twiddle, and runs a
// second 20-point DFT.  Two real frames are packed as one complex signal (re = frame A,
// im = frame B) and separated by Hermitian symmetry.  Host+device so the math is unit-tested on
// the CPU (wb_selftest_fft400) before it ever meets a GPU.
// Replaces rustfft's 400-point plan (reference call sites /root/reference/src/main.rs:440-441,473).
#pragma once
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
WB_HD c32 cadd(c32 a, c32 b) { return {a.x + b.x, a.y + b.y}; }
WB_HD c32 csub(c32 a, c32 b) { return {a.x - b.x, a.y - b.y}; }
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

// W_20^j = exp(-2*pi*i*j/20), j = 0..12 (largest b*c used is 4*3)
WB_HD c32 w20(int j) {
    const float C[13] = {1.0f, 0.95105651629515357f, 0.80901699437494742f, 0.58778525229247313f,
                         0.30901699437494742f, 0.0f, -0.30901699437494742f, -0.58778525229247313f,
                         -0.80901699437494742f, -0.95105651629515357f, -1.0f, -0.95105651629515357f,
                         -0.80901699437494742f};
    const float S[13] = {0.0f, 0.30901699437494742f, 0.58778525229247313f, 0.80901699437494742f,
                         0.95105651629515357f, 1.0f, 0.95105651629515357f, 0.80901699437494742f,
                         0.58778525229247313f, 0.30901699437494742f, 0.0f, -0.30901699437494742f,
                         -0.58778525229247313f};
    return {C[j], -S[j]};
}

// In-place forward 20-point DFT. Input v[n] (natural order), output v[k] (natural order).
WB_HD void dft20(c32 (&v)[20]) {
    // n = 5a + b ; k = c + 4e
    c32 u[5][4];
WB_UNROLL
    for (int b = 0; b < 5; ++b) {
        c32 a0 = v[b], a1 = v[5 + b], a2 = v[10 + b], a3 = v[15 + b];
        dft4(a0, a1, a2, a3);
        u[b][0] = a0;
        u[b][1] = (b == 0) ? a1 : cmul(a1, w20(b * 1));
        u[b][2] = (b == 0) ? a2 : cmul(a2, w20(b * 2));
        u[b][3] = (b == 0) ? a3 : cmul(a3, w20(b * 3));
    }
WB_UNROLL
    for (int c = 0; c < 4; ++c) {
        c32 y0 = u[0][c], y1 = u[1][c], y2 = u[2][c], y3 = u[3][c], y4 = u[4][c];
        dft5(y0, y1, y2, y3, y4);
        v[c] = y0;
        v[c + 4] = y1;
        v[c + 8] = y2;
        v[c + 12] = y3;
        v[c + 16] = y4;
    }
}
