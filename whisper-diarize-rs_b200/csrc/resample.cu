// resample.cu — rational polyphase resampler to 16 kHz mono int16 (north-star piece 1: "16 kHz PCM decode/resample").
//
// The reference refuses anything but 16 kHz / mono / 16-bit (`audio::read_wav`, reference src/audio.rs:9-20) and leaves the
// conversion to the caller (ffmpeg in the examples).  This entry point takes that step onto the device so that a WAV at any
// common rate reaches the log-mel kernel without a host pass.  There is no reference arithmetic to match, so the definition is
// the textbook one (the same one scipy.signal.resample_poly implements, which the oracle is cross-checked against):
//
//   up / down = 16000 / rate (reduced),  half = 10 * max(up, down),  fc = 1 / max(up, down)
//   g[k] = sinc(fc (k - half)) * kaiser_beta5(k),  k = 0 .. 2 half;    h = up * g / sum(g)          (unit DC gain)
//   y[m] = sum_j x[j] * h[m down + half - j up]                        (x = 0 outside the buffer),  m < ceil(n up / down)
//
// with x the channel mean of the interleaved input.  One CTA = a tile of 1024 output samples: the input span it needs and the
// whole tap table are staged in shared memory once (coalesced 16-bit loads, decoded to fp32), every thread then walks its own
// polyphase branch (stride `up` through the taps: conflict-free because gcd(up, down) = 1).  HBM-bound by construction:
// 2 B x channels in + 2 B (+ 4 B) out per sample, each read once apart from the 2 half / up tile halo.
#include <math.h>
#include <vector>
#include "common.cuh"

namespace wdr {

constexpr int kRsTile = 1024;
constexpr int kRsThreads = 256;
constexpr int kRsMaxTaps = 40001;

__global__ void __launch_bounds__(kRsThreads)
resample_kernel(const int16_t* __restrict__ in, int64_t n_in, int channels, int up, int down, int half, const float* __restrict__ taps_g,
                int n_taps, int64_t n_out, int16_t* __restrict__ out_i16, float* __restrict__ out_f32) {
    extern __shared__ __align__(16) float rs_smem[];
    float* taps = rs_smem;
    float* xs = rs_smem + n_taps;
    for (int i = threadIdx.x; i < n_taps; i += kRsThreads) taps[i] = taps_g[i];
    const float inv_ch = 1.0f / (float)channels;
    for (int64_t m0 = (int64_t)blockIdx.x * kRsTile; m0 < n_out; m0 += (int64_t)gridDim.x * kRsTile) {
        const int64_t m1 = min(m0 + (int64_t)kRsTile, n_out);  // exclusive
        // input span of the tile: j_lo = ceil((m0 down - half) / up), j_hi = floor(((m1 - 1) down + half) / up)
        const int64_t a = m0 * down - half;
        const int64_t j_lo = a >= 0 ? (a + up - 1) / up : -((-a) / up);
        const int64_t j_hi = ((m1 - 1) * down + half) / up;
        const int span = (int)(j_hi - j_lo + 1);
        __syncthreads();  // previous tile's readers are done with xs (and the taps are in place)
        for (int i = threadIdx.x; i < span; i += kRsThreads) {
            const int64_t j = j_lo + i;
            float v = 0.0f;
            if (j >= 0 && j < n_in) {
                if (channels == 1) v = (float)in[j];
                else {
                    for (int c = 0; c < channels; c++) v += (float)in[j * channels + c];
                    v *= inv_ch;
                }
            }
            xs[i] = v;
        }
        __syncthreads();
        for (int64_t m = m0 + threadIdx.x; m < m1; m += kRsThreads) {
            const int64_t idx = m * down + half;
            int64_t j = idx / up;             // newest input sample under the filter
            int k = (int)(idx - j * up);      // its tap; older samples sit `up` taps further
            float acc = 0.0f;
            const float* xp = xs + (j - j_lo);
            for (; k < n_taps; k += up, xp--) acc = fmaf(*xp, taps[k], acc);
            if (out_f32) out_f32[m] = acc * (1.0f / 32768.0f);
            if (out_i16) out_i16[m] = (int16_t)__float2int_rn(fminf(fmaxf(acc, -32768.0f), 32767.0f));
        }
    }
}

static double bessel_i0(double x) {
    double sum = 1.0, term = 1.0;
    const double q = 0.25 * x * x;
    for (int k = 1; k < 200; k++) {
        term *= q / ((double)k * (double)k);
        sum += term;
        if (term < 1e-18 * sum) break;
    }
    return sum;
}

static int gcd_i(int a, int b) {
    while (b) { const int t = a % b; a = b; b = t; }
    return a;
}

// taps in double, rounded to fp32 once (the oracle rounds the same way)
static void make_taps(int up, int down, std::vector<float>& h, int* half_out) {
    const int half = 10 * (up > down ? up : down);
    const double fc = 1.0 / (double)(up > down ? up : down);
    const int n = 2 * half + 1;
    std::vector<double> g(n);
    const double i0b = bessel_i0(5.0);
    double sum = 0.0;
    for (int k = 0; k < n; k++) {
        const double t = (double)(k - half);
        const double x = fc * t;
        const double s = t == 0.0 ? 1.0 : sin(M_PI * x) / (M_PI * x);
        const double r = t / (double)half;
        const double w = bessel_i0(5.0 * sqrt(fmax(0.0, 1.0 - r * r))) / i0b;
        g[k] = s * w;
        sum += g[k];
    }
    h.resize(n);
    for (int k = 0; k < n; k++) h[k] = (float)((double)up * g[k] / sum);
    *half_out = half;
}

}  // namespace wdr

using namespace wdr;

extern "C" int64_t wdr_resample_n_out(int64_t n_frames, int sample_rate) {
    if (n_frames < 0 || sample_rate <= 0) return -1;
    const int g = gcd_i(16000, sample_rate);
    const int64_t up = 16000 / g, down = sample_rate / g;
    return (n_frames * up + down - 1) / down;
}

extern "C" int wdr_resample_i16(const int16_t* pcm, int64_t n_frames, int channels, int sample_rate, int16_t* out_i16, float* out_f32,
                                int64_t out_cap, int64_t* n_out) {
    clear_error();
    WDR_REQUIRE(n_frames >= 0 && channels >= 1 && channels <= 8 && sample_rate > 0 && (n_frames == 0 || pcm) && n_out, "bad arguments");
    const int g = gcd_i(16000, sample_rate);
    const int up = 16000 / g, down = sample_rate / g;
    const int64_t n = (n_frames * up + down - 1) / down;
    *n_out = n;
    if (2 * 10 * (up > down ? up : down) + 1 > kRsMaxTaps) {
        set_error("sample rate %d Hz needs a %d-tap filter (limit %d): not a supported rate", sample_rate, 20 * (up > down ? up : down) + 1, kRsMaxTaps);
        return WDR_ERR_UNSUPPORTED;
    }
    WDR_REQUIRE(n <= out_cap && (n == 0 || out_i16 || out_f32), "output buffer too small");
    int rc = ensure_device(-1);
    if (rc != WDR_OK) return rc;
    if (n == 0) return WDR_OK;
    std::vector<float> h;
    int half = 0;
    make_taps(up, down, h, &half);
    const int n_taps = (int)h.size();
    const int span_max = (int)(((int64_t)(kRsTile - 1) * down + 2 * half) / up + 2);
    const size_t smem = sizeof(float) * ((size_t)n_taps + span_max);
    WDR_REQUIRE(smem <= 200 * 1024, "sample rate ratio needs more shared memory than a CTA has");
    WDR_CUDA_TRY(cudaFuncSetAttribute(resample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DevBuf<int16_t> d_in, d_o16;
    DevBuf<float> d_h, d_o32;
    WDR_CUDA_TRY(d_in.alloc((size_t)n_frames * channels));
    WDR_CUDA_TRY(d_h.alloc(n_taps));
    if (out_i16) WDR_CUDA_TRY(d_o16.alloc(n));
    if (out_f32) WDR_CUDA_TRY(d_o32.alloc(n));
    WDR_CUDA_TRY(cudaMemcpy(d_in.p, pcm, sizeof(int16_t) * (size_t)n_frames * channels, cudaMemcpyHostToDevice));
    WDR_CUDA_TRY(cudaMemcpy(d_h.p, h.data(), sizeof(float) * n_taps, cudaMemcpyHostToDevice));
    const int64_t tiles = (n + kRsTile - 1) / kRsTile;
    const int grid = (int)(tiles < 148 * 2 ? tiles : 148 * 2);
    resample_kernel<<<grid, kRsThreads, smem>>>(d_in.p, n_frames, channels, up, down, half, d_h.p, n_taps, n, out_i16 ? d_o16.p : nullptr,
                                               out_f32 ? d_o32.p : nullptr);
    WDR_LAUNCH_CHECK();
    if (out_i16) WDR_CUDA_TRY(cudaMemcpy(out_i16, d_o16.p, sizeof(int16_t) * n, cudaMemcpyDeviceToHost));
    if (out_f32) WDR_CUDA_TRY(cudaMemcpy(out_f32, d_o32.p, sizeof(float) * n, cudaMemcpyDeviceToHost));
    return WDR_OK;
}
