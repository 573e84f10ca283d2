// mel_core.cuh — the per-task arithmetic of the log-mel kernel (whisper.cpp log_mel_spectrogram,
// reached from reference src/transcribe.rs:389).  Two real frames are packed into one complex 400-point
// FFT (400 = 16 x 25, two shared-memory passes); the three tasks below are what one thread executes per
// pass.  __host__ __device__ so the CPU test harness (tests/host_harness.cpp) runs the identical code.
#pragma once
#include "fft_small.cuh"

#define MEL_NFFT 400
#define MEL_HOP 160
#define MEL_NBINS 201
#define MEL_FRAMES_PER_CTA 32
#define MEL_PAIRS_PER_CTA (MEL_FRAMES_PER_CTA / 2)
#define MEL_TILE_SAMPLES ((MEL_FRAMES_PER_CTA - 1) * MEL_HOP + MEL_NFFT) /* 5360 */
#define MEL_ZPITCH 409 /* complex elements per pair in shared memory; == 25 (mod 16) keeps 64-bit accesses conflict-free */
#define MEL_PPITCH 33  /* power-spectrum row pitch (frames + 1) */

// Pass 1, task (pair p, b in [0,25)): windowed samples n = 25a + b of frames 2p (real part) and 2p+1
// (imaginary part), 16-point DFT over a, twiddle W400^(b*k1), store t[k1][b].
WDR_HD void mel_pass1_task(const float* tile, const float* hann, const cpx* tw400, cpx* zbuf, int p, int b) {
    cpx x[16];
    const float* f0 = tile + (2 * p) * MEL_HOP;
    const float* f1 = f0 + MEL_HOP;
#pragma unroll
    for (int a = 0; a < 16; a++) {
        const int n = 25 * a + b;
        const float w = hann[n];
        x[a] = cmake(w * f0[n], w * f1[n]);
    }
    dft16(x);
    cpx* z = zbuf + p * MEL_ZPITCH;
    z[b] = x[0];
#pragma unroll
    for (int k1 = 1; k1 < 16; k1++) z[k1 * 25 + b] = cmul(x[k1], tw400[b * k1]);
}

// Pass 2, task (pair p, k1 in [0,16)): 25-point DFT over b, in place: slot [k1*25 + k2] <- X[k1 + 16*k2].
WDR_HD void mel_pass2_task(const cpx* tw25, cpx* zbuf, int p, int k1) {
    cpx x[25];
    cpx* z = zbuf + p * MEL_ZPITCH + k1 * 25;
#pragma unroll
    for (int b = 0; b < 25; b++) x[b] = z[b];
    dft25(x, tw25);
#pragma unroll
    for (int b = 0; b < 25; b++) z[b] = x[b];
}

// Pass 3, task (pair p, bin k in [0,201)): split the packed spectrum into the two real-input spectra and
// store their power: F0 = (Z[k] + conj Z[N-k]) / 2, F1 = (Z[k] - conj Z[N-k]) / (2i).
WDR_HD void mel_pass3_task(const cpx* zbuf, float* pbuf, int p, int k) {
    const cpx* z = zbuf + p * MEL_ZPITCH;
    const int kr = (MEL_NFFT - k) % MEL_NFFT;
    const cpx za = z[(k & 15) * 25 + (k >> 4)];
    const cpx zb = z[(kr & 15) * 25 + (kr >> 4)];
    const float ar = za.re + zb.re, ai = za.im - zb.im;
    const float br = za.im + zb.im, bi = zb.re - za.re;
    pbuf[k * MEL_PPITCH + 2 * p] = 0.25f * (ar * ar + ai * ai);
    pbuf[k * MEL_PPITCH + 2 * p + 1] = 0.25f * (br * br + bi * bi);
}
