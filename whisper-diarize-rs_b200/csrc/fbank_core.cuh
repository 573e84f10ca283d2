// fbank_core.cuh — per-task arithmetic of the Kaldi fbank kernel (kaldi-native-fbank via knf-rs
// compute_fbank inside pyannote-rs EmbeddingExtractor::compute, reference src/transcribe.rs:466).
// Two real frames are packed into one complex FFT-512 = 16 x 32 (32 = 2 x 16).  DC removal,
// pre-emphasis 0.97 and the povey window are fused into the first pass.  __host__ __device__ so the CPU
// harness runs the same code.
#pragma once
#include "fft_small.cuh"

#define FB_FLEN 400
#define FB_SHIFT 160
#define FB_NFFT 512
#define FB_NBINS 256 /* power bins the mel banks see (Nyquist excluded, as in Kaldi) */
#define FB_FRAMES_PER_CTA 32
#define FB_PAIRS_PER_CTA (FB_FRAMES_PER_CTA / 2)
#define FB_TILE_SAMPLES ((FB_FRAMES_PER_CTA - 1) * FB_SHIFT + FB_FLEN) /* 5360 */
#define FB_ROWPITCH 33                 /* complex elements per k1 row (32 + 1 pad) */
#define FB_ZPITCH (16 * FB_ROWPITCH)   /* complex elements per pair */
#define FB_PPITCH 33

// pre-processed, windowed sample n of a frame starting at `f` (raw int16-scale floats), frame mean `mu`
WDR_HD float fb_sample(const float* f, float mu, const float* window, int n) {
    if (n >= FB_FLEN) return 0.0f;
    const float d = f[n] - mu;
    const float dp = (n == 0) ? d : (f[n - 1] - mu);
    return (d - 0.97f * dp) * window[n];
}

// Pass 1, task (pair p, b in [0,32)): samples n = 32a + b, 16-point DFT over a, twiddle W512^(b*k1).
WDR_HD void fb_pass1_task(const float* tile, const float* mean, const float* window, const cpx* tw512, cpx* zbuf, int p, int b) {
    cpx x[16];
    const float* f0 = tile + (2 * p) * FB_SHIFT;
    const float* f1 = f0 + FB_SHIFT;
    const float m0 = mean[2 * p], m1 = mean[2 * p + 1];
#pragma unroll
    for (int a = 0; a < 16; a++) {
        const int n = 32 * a + b;
        x[a] = cmake(fb_sample(f0, m0, window, n), fb_sample(f1, m1, window, n));
    }
    dft16(x);
    cpx* z = zbuf + p * FB_ZPITCH;
    z[b] = x[0];
#pragma unroll
    for (int k1 = 1; k1 < 16; k1++) z[k1 * FB_ROWPITCH + b] = cmul(x[k1], tw512[b * k1]);
}

// Forward 32-point DFT = 2 x 16, in place, natural order.  tw32[j] = exp(-2 pi i j / 32), j < 16.
WDR_HD void dft32(cpx* x, const cpx* tw32) {
    // b = 16*b1 + b0, k = q1 + 2*q2: DFT2 over b1, twiddle W32^(b0*q1), DFT16 over b0
    cpx e[16], o[16];
#pragma unroll
    for (int b0 = 0; b0 < 16; b0++) {
        e[b0] = cadd(x[b0], x[16 + b0]);
        o[b0] = cmul(csub(x[b0], x[16 + b0]), tw32[b0]);
    }
    dft16(e);
    dft16(o);
#pragma unroll
    for (int q2 = 0; q2 < 16; q2++) {
        x[2 * q2] = e[q2];
        x[2 * q2 + 1] = o[q2];
    }
}

// Pass 2, task (pair p, k1 in [0,16)): 32-point DFT over b in place: slot [k1][k2] <- X[k1 + 16*k2].
WDR_HD void fb_pass2_task(const cpx* tw32, cpx* zbuf, int p, int k1) {
    cpx x[32];
    cpx* z = zbuf + p * FB_ZPITCH + k1 * FB_ROWPITCH;
#pragma unroll
    for (int b = 0; b < 32; b++) x[b] = z[b];
    dft32(x, tw32);
#pragma unroll
    for (int b = 0; b < 32; b++) z[b] = x[b];
}

// Pass 3, task (pair p, bin k in [0,256)): power spectra of the two packed real frames.
WDR_HD void fb_pass3_task(const cpx* zbuf, float* pbuf, int p, int k) {
    const cpx* z = zbuf + p * FB_ZPITCH;
    const int kr = (FB_NFFT - k) % FB_NFFT;
    const cpx za = z[(k & 15) * FB_ROWPITCH + (k >> 4)];
    const cpx zb = z[(kr & 15) * FB_ROWPITCH + (kr >> 4)];
    const float ar = za.re + zb.re, ai = za.im - zb.im;
    const float br = za.im + zb.im, bi = zb.re - za.re;
    pbuf[k * FB_PPITCH + 2 * p] = 0.25f * (ar * ar + ai * ai);
    pbuf[k * FB_PPITCH + 2 * p + 1] = 0.25f * (br * br + bi * bi);
}
