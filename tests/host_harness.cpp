// host_harness.cpp — compiles the __host__ __device__ arithmetic of the CUDA kernels (fft_small.cuh,
// mel_core.cuh, fbank_core.cuh) with g++ so the CPU test-suite can check it against numpy without a GPU.
// Test infrastructure only; built by tests/conftest.py into tests/_build/libharness.so.
#include <cmath>
#include <cstring>
#include <vector>
#include "mel_core.cuh"

extern "C" {

void harness_dft16(float* io /* 16 complex interleaved */) { dft16(reinterpret_cast<cpx*>(io)); }

void harness_dft25(float* io) {
    cpx tw[25];
    for (int j = 0; j < 25; j++) tw[j] = cmake((float)cos(2.0 * M_PI * j / 25), (float)-sin(2.0 * M_PI * j / 25));
    dft25(reinterpret_cast<cpx*>(io), tw);
}

// tile: MEL_TILE_SAMPLES floats.  power: [201][32] floats (frame fastest).
void harness_mel_power(const float* tile, float* power) {
    std::vector<float> hann(MEL_NFFT);
    std::vector<cpx> tw400(MEL_NFFT), tw25(25);
    for (int i = 0; i < MEL_NFFT; i++) {
        hann[i] = (float)(0.5 * (1.0 - cos(2.0 * M_PI * i / MEL_NFFT)));
        tw400[i] = cmake((float)cos(2.0 * M_PI * i / MEL_NFFT), (float)-sin(2.0 * M_PI * i / MEL_NFFT));
    }
    for (int j = 0; j < 25; j++) tw25[j] = cmake((float)cos(2.0 * M_PI * j / 25), (float)-sin(2.0 * M_PI * j / 25));
    std::vector<cpx> zbuf(MEL_PAIRS_PER_CTA * MEL_ZPITCH);
    std::vector<float> pbuf(MEL_NBINS * MEL_PPITCH);
    for (int t = 0; t < MEL_PAIRS_PER_CTA * 25; t++) mel_pass1_task(tile, hann.data(), tw400.data(), zbuf.data(), t / 25, t % 25);
    for (int t = 0; t < MEL_PAIRS_PER_CTA * 16; t++) mel_pass2_task(tw25.data(), zbuf.data(), t / 16, t % 16);
    for (int t = 0; t < MEL_PAIRS_PER_CTA * MEL_NBINS; t++) mel_pass3_task(zbuf.data(), pbuf.data(), t / MEL_NBINS, t % MEL_NBINS);
    for (int k = 0; k < MEL_NBINS; k++)
        for (int f = 0; f < MEL_FRAMES_PER_CTA; f++) power[k * MEL_FRAMES_PER_CTA + f] = pbuf[k * MEL_PPITCH + f];
}

}  // extern "C"

#include "fbank_core.cuh"
extern "C" {
// tile: FB_TILE_SAMPLES int16-scale floats.  power: [256][32].
void harness_fbank_power(const float* tile, float* power) {
    std::vector<float> window(FB_FLEN), mean(FB_FRAMES_PER_CTA);
    std::vector<cpx> tw512(FB_NFFT), tw32(16);
    for (int i = 0; i < FB_FLEN; i++) window[i] = (float)pow(0.5 - 0.5 * cos(2.0 * M_PI / (FB_FLEN - 1) * i), 0.85);
    for (int i = 0; i < FB_NFFT; i++) tw512[i] = cmake((float)cos(2.0 * M_PI * i / FB_NFFT), (float)-sin(2.0 * M_PI * i / FB_NFFT));
    for (int j = 0; j < 16; j++) tw32[j] = cmake((float)cos(2.0 * M_PI * j / 32), (float)-sin(2.0 * M_PI * j / 32));
    for (int f = 0; f < FB_FRAMES_PER_CTA; f++) {
        float s = 0;
        for (int i = 0; i < FB_FLEN; i++) s += tile[f * FB_SHIFT + i];
        mean[f] = s / FB_FLEN;
    }
    std::vector<cpx> zbuf(FB_PAIRS_PER_CTA * FB_ZPITCH);
    std::vector<float> pbuf(FB_NBINS * FB_PPITCH);
    for (int t = 0; t < FB_PAIRS_PER_CTA * 32; t++) fb_pass1_task(tile, mean.data(), window.data(), tw512.data(), zbuf.data(), t / 32, t % 32);
    for (int t = 0; t < FB_PAIRS_PER_CTA * 16; t++) fb_pass2_task(tw32.data(), zbuf.data(), t / 16, t % 16);
    for (int t = 0; t < FB_PAIRS_PER_CTA * FB_NBINS; t++) fb_pass3_task(zbuf.data(), pbuf.data(), t / FB_NBINS, t % FB_NBINS);
    for (int k = 0; k < FB_NBINS; k++)
        for (int f = 0; f < FB_FRAMES_PER_CTA; f++) power[k * FB_FRAMES_PER_CTA + f] = pbuf[k * FB_PPITCH + f];
}
}
