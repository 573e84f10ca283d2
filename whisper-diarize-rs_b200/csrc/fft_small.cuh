// fft_small.cuh — register-resident small DFTs (4, 5, 16, 25 points) used by the log-mel and Kaldi
// fbank kernels.  All functions are __host__ __device__ so that tests/test_fft_host.py can compile the
// same source with g++ and check it against numpy on the CPU box before any GPU time is spent.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define WDR_HD __host__ __device__ __forceinline__
#else
#define WDR_HD inline
#endif

struct cpx {
    float re, im;
};

WDR_HD cpx cmake(float re, float im) { cpx c; c.re = re; c.im = im; return c; }
WDR_HD cpx cadd(cpx a, cpx b) { return cmake(a.re + b.re, a.im + b.im); }
WDR_HD cpx csub(cpx a, cpx b) { return cmake(a.re - b.re, a.im - b.im); }
WDR_HD cpx cmul(cpx a, cpx b) { return cmake(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
// multiply by -i : (re, im) -> (im, -re)
WDR_HD cpx cmul_mi(cpx a) { return cmake(a.im, -a.re); }

// Forward 4-point DFT, in place: y[k] = sum_n x[n] exp(-2 pi i n k / 4)
WDR_HD void dft4(cpx& x0, cpx& x1, cpx& x2, cpx& x3) {
    cpx s02 = cadd(x0, x2), d02 = csub(x0, x2);
    cpx s13 = cadd(x1, x3), d13 = cmul_mi(csub(x1, x3));
    x0 = cadd(s02, s13);
    x1 = cadd(d02, d13);
    x2 = csub(s02, s13);
    x3 = csub(d02, d13);
}

// Forward 5-point DFT, in place.
WDR_HD void dft5(cpx& x0, cpx& x1, cpx& x2, cpx& x3, cpx& x4) {
    const float c1 = 0.30901699437494742f;   // cos(2pi/5)
    const float c2 = -0.80901699437494742f;  // cos(4pi/5)
    const float s1 = 0.95105651629515357f;   // sin(2pi/5)
    const float s2 = 0.58778525229247313f;   // sin(4pi/5)
    cpx a1 = cadd(x1, x4), d1 = csub(x1, x4);
    cpx a2 = cadd(x2, x3), d2 = csub(x2, x3);
    cpx A1 = cmake(x0.re + c1 * a1.re + c2 * a2.re, x0.im + c1 * a1.im + c2 * a2.im);
    cpx A2 = cmake(x0.re + c2 * a1.re + c1 * a2.re, x0.im + c2 * a1.im + c1 * a2.im);
    cpx B1 = cmake(s1 * d1.re + s2 * d2.re, s1 * d1.im + s2 * d2.im);
    cpx B2 = cmake(s2 * d1.re - s1 * d2.re, s2 * d1.im - s1 * d2.im);
    cpx y0 = cmake(x0.re + a1.re + a2.re, x0.im + a1.im + a2.im);
    // y1 = A1 - i B1, y4 = A1 + i B1, y2 = A2 - i B2, y3 = A2 + i B2
    x0 = y0;
    x1 = cmake(A1.re + B1.im, A1.im - B1.re);
    x4 = cmake(A1.re - B1.im, A1.im + B1.re);
    x2 = cmake(A2.re + B2.im, A2.im - B2.re);
    x3 = cmake(A2.re - B2.im, A2.im + B2.re);
}

// Forward 16-point DFT (4 x 4 Cooley-Tukey), in place, natural order in and out.
//   a = 4*a1 + a0, k = k1 + 4*k2
WDR_HD void dft16(cpx* x) {
    const float C1 = 0.92387953251128674f;  // cos(pi/8)
    const float S1 = 0.38268343236508977f;  // sin(pi/8)
    const float R = 0.70710678118654752f;   // cos(pi/4)
    // step 1: for each a0, DFT4 over a1 on x[4*a1 + a0] -> t[a0][k1] kept in x[4*k1 + a0]
#pragma unroll
    for (int a0 = 0; a0 < 4; a0++) dft4(x[a0], x[4 + a0], x[8 + a0], x[12 + a0]);
    // step 2: twiddle W16^(a0*k1), W16 = exp(-2 pi i/16)
    // k1 = 1: a0 = 1,2,3 -> W^1, W^2, W^3
    x[4 + 1] = cmul(x[4 + 1], cmake(C1, -S1));
    x[4 + 2] = cmul(x[4 + 2], cmake(R, -R));
    x[4 + 3] = cmul(x[4 + 3], cmake(S1, -C1));
    // k1 = 2: W^2, W^4, W^6
    x[8 + 1] = cmul(x[8 + 1], cmake(R, -R));
    x[8 + 2] = cmul_mi(x[8 + 2]);
    x[8 + 3] = cmul(x[8 + 3], cmake(-R, -R));
    // k1 = 3: W^3, W^6, W^9
    x[12 + 1] = cmul(x[12 + 1], cmake(S1, -C1));
    x[12 + 2] = cmul(x[12 + 2], cmake(-R, -R));
    x[12 + 3] = cmul(x[12 + 3], cmake(-C1, S1));
    // step 3: for each k1, DFT4 over a0 -> Y[k1 + 4*k2] lands in x[4*k1 + k2]
#pragma unroll
    for (int k1 = 0; k1 < 4; k1++) dft4(x[4 * k1], x[4 * k1 + 1], x[4 * k1 + 2], x[4 * k1 + 3]);
    // transpose to natural order: Y[k1 + 4*k2] currently at x[4*k1 + k2]
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = i + 1; j < 4; j++) {
            cpx t = x[4 * i + j];
            x[4 * i + j] = x[4 * j + i];
            x[4 * j + i] = t;
        }
}

// Forward 25-point DFT (5 x 5), in place, natural order in and out.  tw25[j] = exp(-2 pi i j / 25),
// j = 0..16 (products b0*q1 <= 16).
WDR_HD void dft25(cpx* x, const cpx* tw25) {
    // b = 5*b1 + b0, k = q1 + 5*q2.  step 1: for each b0, DFT5 over b1 -> t[b0][q1] in x[5*q1 + b0]
#pragma unroll
    for (int b0 = 0; b0 < 5; b0++) dft5(x[b0], x[5 + b0], x[10 + b0], x[15 + b0], x[20 + b0]);
#pragma unroll
    for (int q1 = 1; q1 < 5; q1++)
#pragma unroll
        for (int b0 = 1; b0 < 5; b0++) x[5 * q1 + b0] = cmul(x[5 * q1 + b0], tw25[b0 * q1]);
#pragma unroll
    for (int q1 = 0; q1 < 5; q1++) dft5(x[5 * q1], x[5 * q1 + 1], x[5 * q1 + 2], x[5 * q1 + 3], x[5 * q1 + 4]);
    // Y[q1 + 5*q2] is at x[5*q1 + q2]: transpose
#pragma unroll
    for (int i = 0; i < 5; i++)
#pragma unroll
        for (int j = i + 1; j < 5; j++) {
            cpx t = x[5 * i + j];
            x[5 * i + j] = x[5 * j + i];
            x[5 * j + i] = t;
        }
}

// Forward 2-point butterfly.
WDR_HD void dft2(cpx& a, cpx& b) {
    cpx s = cadd(a, b), d = csub(a, b);
    a = s;
    b = d;
}
