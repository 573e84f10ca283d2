// fbank.cu — Kaldi-compatible fbank features on the device.
//
// Replaces kaldi-native-fbank as driven by knf-rs `compute_fbank` inside pyannote-rs
// `EmbeddingExtractor::compute` (reference call site src/transcribe.rs:466; options in SURVEY A.9):
// 25 ms / 10 ms frames, snip_edges, dither 0, remove_dc_offset, preemph 0.97, povey window, 512-point
// power spectrum, HTK-mel triangular banks 20 Hz..Nyquist, log(max(e, FLT_EPSILON)), then pyannote-rs'
// per-utterance column mean subtraction.
//
// Same shape as the log-mel kernel: one CTA = 32 frames of one segment staged once in shared memory, two
// frames per complex FFT-512 (16 x 32), bin-major power spectrum, sparse banks with lane = frame; the
// [32][n_bins] result block is contiguous in the frame-major output, so it is staged and written as
// coalesced rows.  The kernel also signals get_signal_energy (whisper.cpp) used by the token-timestamp
// heuristic.
#include <float.h>
#include <math.h>
#include <vector>
#include "common.cuh"
#include "fbank_core.cuh"

namespace wdr {

constexpr int kFbThreads = 256;
constexpr int kFbMaxBins = 128;
constexpr size_t kFbSmemBytes = FB_TILE_SAMPLES * sizeof(float) + FB_FLEN * sizeof(float) + FB_NFFT * sizeof(cpx) +
                                FB_PAIRS_PER_CTA * FB_ZPITCH * sizeof(cpx) + FB_NBINS * FB_PPITCH * sizeof(float) +
                                FB_FRAMES_PER_CTA * sizeof(float);

__constant__ cpx c_tw32[16];

struct FbankTables {
    float* d_window = nullptr;
    cpx* d_tw512 = nullptr;
    float* d_fw = nullptr;
    int4* d_frow = nullptr;
    int n_bins = 0;
    int device = -1;
};
static FbankTables g_fb[16];  // per device

// grid: (cta index within the launch); cta_map[blockIdx.x] = {segment, first frame}
__global__ void __launch_bounds__(kFbThreads, 2)
fbank_kernel(const int16_t* __restrict__ pcm, const int64_t* __restrict__ seg_offset, const int64_t* __restrict__ seg_end /* null: segment s ends where s + 1 starts */,
             const int64_t* __restrict__ feat_offset, const int2* __restrict__ cta_map, const float* __restrict__ window_g, const cpx* __restrict__ tw512_g,
             const float* __restrict__ fw, const int4* __restrict__ frow, int n_bins, float* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* tile = reinterpret_cast<float*>(smem_raw);
    float* window = tile + FB_TILE_SAMPLES;
    cpx* tw512 = reinterpret_cast<cpx*>(window + FB_FLEN);
    cpx* zbuf = tw512 + FB_NFFT;
    float* pbuf = reinterpret_cast<float*>(zbuf + FB_PAIRS_PER_CTA * FB_ZPITCH);
    float* mean = pbuf + FB_NBINS * FB_PPITCH;
    float* stage = tile;  // the sample tile is dead after pass 1: reuse it for the [32][n_bins] output block

    const int2 cm = cta_map[blockIdx.x];
    const int seg = cm.x, frame0 = cm.y;
    const int64_t s_begin = seg_offset[seg];
    const int n = (int)((seg_end ? seg_end[seg] : seg_offset[seg + 1]) - s_begin);
    const int T = (n < FB_FLEN) ? 0 : 1 + (n - FB_FLEN) / FB_SHIFT;
    const int frames_here = min(FB_FRAMES_PER_CTA, T - frame0);
    if (frames_here <= 0) return;
    const int16_t* x = pcm + s_begin + (int64_t)frame0 * FB_SHIFT;
    const int avail = n - frame0 * FB_SHIFT;  // samples available from the tile start
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int j = tid; j < FB_TILE_SAMPLES; j += kFbThreads) tile[j] = (j < avail) ? (float)x[j] : 0.0f;
    for (int i = tid; i < FB_FLEN; i += kFbThreads) window[i] = window_g[i];
    for (int i = tid; i < FB_NFFT; i += kFbThreads) tw512[i] = tw512_g[i];
    __syncthreads();
    // frame means (remove_dc_offset): warp w sums frames w, w+8, ...
    for (int f = warp; f < FB_FRAMES_PER_CTA; f += kFbThreads / 32) {
        float s = 0.0f;
        for (int i = lane; i < FB_FLEN; i += 32) s += tile[f * FB_SHIFT + i];
        s = warp_sum(s);
        if (lane == 0) mean[f] = s / FB_FLEN;
    }
    __syncthreads();
    for (int t = tid; t < FB_PAIRS_PER_CTA * 32; t += kFbThreads) fb_pass1_task(tile, mean, window, tw512, zbuf, t >> 5, t & 31);
    __syncthreads();
    for (int t = tid; t < FB_PAIRS_PER_CTA * 16; t += kFbThreads) fb_pass2_task(c_tw32, zbuf, t >> 4, t & 15);
    __syncthreads();
    for (int t = tid; t < FB_PAIRS_PER_CTA * FB_NBINS; t += kFbThreads) fb_pass3_task(zbuf, pbuf, t >> 8, t & 255);
    __syncthreads();
    for (int m = warp; m < n_bins; m += kFbThreads / 32) {
        const int4 r = __ldg(&frow[m]);
        const float* w = fw + r.z;
        const float* p = pbuf + r.x * FB_PPITCH + lane;
        float acc = 0.0f;
        for (int k = 0; k < r.y; k++) acc = fmaf(__ldg(&w[k]), p[k * FB_PPITCH], acc);
        stage[lane * n_bins + m] = logf(fmaxf(acc, FLT_EPSILON));
    }
    __syncthreads();
    float* o = out + (feat_offset[seg] + frame0) * n_bins;
    for (int i = tid; i < frames_here * n_bins; i += kFbThreads) o[i] = stage[i];
}

// per-segment column mean subtraction: one CTA per segment, 4 row groups x n_bins columns
__global__ void fbank_cmn_kernel(const int64_t* __restrict__ feat_offset, int n_bins, float* __restrict__ feats) {
    extern __shared__ float part[];  // [groups][n_bins]
    const int seg = blockIdx.x;
    const int64_t f0 = feat_offset[seg];
    const int T = (int)(feat_offset[seg + 1] - f0);
    if (T <= 0) return;
    const int groups = blockDim.x / n_bins;
    const int g = threadIdx.x / n_bins, b = threadIdx.x % n_bins;
    float* base = feats + f0 * n_bins;
    double s = 0.0;
    if (g < groups)
        for (int t = g; t < T; t += groups) s += (double)base[(int64_t)t * n_bins + b];
    if (g < groups) part[g * n_bins + b] = (float)s;
    __syncthreads();
    if (g < groups) {
        float tot = 0.0f;
        for (int q = 0; q < groups; q++) tot += part[q * n_bins + b];
        const float mu = tot / (float)T;
        for (int t = g; t < T; t += groups) base[(int64_t)t * n_bins + b] -= mu;
    }
}

__global__ void signal_energy_kernel(const float* __restrict__ x, int n, int hw, float* __restrict__ out) {
    const int stride = gridDim.x * blockDim.x;
    const float den = (float)(2 * hw + 1);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float s = 0.0f;
        const int lo = max(0, i - hw), hi = min(n - 1, i + hw);
        for (int j = lo; j <= hi; j++) s += fabsf(__ldg(&x[j]));
        out[i] = s / den;
    }
}

static inline float mel_scale_f(float f) { return 1127.0f * logf(1.0f + f / 700.0f); }

static int fbank_tables(int n_bins, FbankTables** out_tab) {
    int dev = 0;
    WDR_CUDA_TRY(cudaGetDevice(&dev));
    WDR_REQUIRE(dev >= 0 && dev < 16, "device index out of range");
    FbankTables& t = g_fb[dev];
    if (t.device == dev && t.n_bins == n_bins) { *out_tab = &t; return WDR_OK; }
    if (t.d_window) { cudaFree(t.d_window); cudaFree(t.d_tw512); cudaFree(t.d_fw); cudaFree(t.d_frow); t = FbankTables(); }
    std::vector<float> window(FB_FLEN);
    std::vector<cpx> tw(FB_NFFT);
    cpx tw32[16];
    const double a = 2.0 * M_PI / (FB_FLEN - 1);
    for (int i = 0; i < FB_FLEN; i++) window[i] = (float)pow(0.5 - 0.5 * cos(a * i), 0.85);
    for (int i = 0; i < FB_NFFT; i++) { tw[i].re = (float)cos(2.0 * M_PI * i / FB_NFFT); tw[i].im = (float)-sin(2.0 * M_PI * i / FB_NFFT); }
    for (int j = 0; j < 16; j++) { tw32[j].re = (float)cos(2.0 * M_PI * j / 32); tw32[j].im = (float)-sin(2.0 * M_PI * j / 32); }
    // Kaldi MelBanks (float arithmetic, as kaldi-native-fbank): low 20 Hz, high = Nyquist, no VTLN
    const float fft_bin_width = (float)WDR_SAMPLE_RATE / FB_NFFT;
    const float mel_low = mel_scale_f(20.0f), mel_high = mel_scale_f(0.5f * WDR_SAMPLE_RATE);
    const float mel_delta = (mel_high - mel_low) / (n_bins + 1);
    std::vector<int4> rows(n_bins);
    std::vector<float> fw;
    for (int b = 0; b < n_bins; b++) {
        const float left = mel_low + b * mel_delta, center = mel_low + (b + 1) * mel_delta, right = mel_low + (b + 2) * mel_delta;
        int lo = -1, hi = -1;
        std::vector<float> wrow(FB_NBINS, 0.0f);
        for (int i = 0; i < FB_NBINS; i++) {
            const float mel = mel_scale_f(fft_bin_width * i);
            if (mel > left && mel < right) {
                wrow[i] = (mel <= center) ? (mel - left) / (center - left) : (right - mel) / (right - center);
                if (lo < 0) lo = i;
                hi = i;
            }
        }
        const int len = lo >= 0 ? hi - lo + 1 : 0;
        rows[b] = make_int4(lo >= 0 ? lo : 0, len, (int)fw.size(), 0);
        for (int k = 0; k < len; k++) fw.push_back(wrow[lo + k]);
    }
    if (fw.empty()) fw.push_back(0.0f);
    WDR_CUDA_TRY(cudaMalloc(&t.d_window, sizeof(float) * FB_FLEN));
    WDR_CUDA_TRY(cudaMalloc(&t.d_tw512, sizeof(cpx) * FB_NFFT));
    WDR_CUDA_TRY(cudaMalloc(&t.d_fw, sizeof(float) * fw.size()));
    WDR_CUDA_TRY(cudaMalloc(&t.d_frow, sizeof(int4) * n_bins));
    WDR_CUDA_TRY(cudaMemcpy(t.d_window, window.data(), sizeof(float) * FB_FLEN, cudaMemcpyHostToDevice));
    WDR_CUDA_TRY(cudaMemcpy(t.d_tw512, tw.data(), sizeof(cpx) * FB_NFFT, cudaMemcpyHostToDevice));
    WDR_CUDA_TRY(cudaMemcpy(t.d_fw, fw.data(), sizeof(float) * fw.size(), cudaMemcpyHostToDevice));
    WDR_CUDA_TRY(cudaMemcpy(t.d_frow, rows.data(), sizeof(int4) * n_bins, cudaMemcpyHostToDevice));
    WDR_CUDA_TRY(cudaMemcpyToSymbol(c_tw32, tw32, sizeof(tw32)));
    WDR_CUDA_TRY(cudaFuncSetAttribute(fbank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFbSmemBytes));
    t.n_bins = n_bins;
    t.device = dev;
    *out_tab = &t;
    return WDR_OK;
}

// seg_offset_host: n_segments+1 sample offsets (host copy, used to build the CTA map).  seg_end_dev / seg_end_host (both or neither):
// segments that are NOT back to back in pcm — segment s then spans [seg_offset[s], seg_end[s]) and seg_offset_host holds n_segments starts.
int fbank_run(const int16_t* pcm, const int64_t* seg_offset_dev, const int64_t* feat_offset_dev,
              const std::vector<int64_t>& seg_offset_host, int n_bins, int subtract_mean, float* out, cudaStream_t st,
              const int64_t* seg_end_dev, const std::vector<int64_t>* seg_end_host) {
    FbankTables* tab = nullptr;
    int rc = fbank_tables(n_bins, &tab);
    if (rc != WDR_OK) return rc;
    const int n_segments = seg_end_host ? (int)seg_end_host->size() : (int)seg_offset_host.size() - 1;
    std::vector<int2> map;
    for (int s = 0; s < n_segments; s++) {
        const int64_t n = (seg_end_host ? (*seg_end_host)[s] : seg_offset_host[s + 1]) - seg_offset_host[s];
        const int T = n < FB_FLEN ? 0 : (int)(1 + (n - FB_FLEN) / FB_SHIFT);
        for (int f = 0; f < T; f += FB_FRAMES_PER_CTA) map.push_back(make_int2(s, f));
    }
    if (map.empty()) return WDR_OK;
    int2* d_map = nullptr;
    WDR_CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&d_map), sizeof(int2) * map.size(), st));
    WDR_CUDA_TRY(cudaMemcpyAsync(d_map, map.data(), sizeof(int2) * map.size(), cudaMemcpyHostToDevice, st));
    fbank_kernel<<<(unsigned)map.size(), kFbThreads, kFbSmemBytes, st>>>(pcm, seg_offset_dev, seg_end_dev, feat_offset_dev, d_map, tab->d_window,
                                                                         tab->d_tw512, tab->d_fw, tab->d_frow, n_bins, out);
    WDR_LAUNCH_CHECK();
    WDR_CUDA_TRY(cudaFreeAsync(d_map, st));
    if (subtract_mean) {
        const int groups = 4;
        fbank_cmn_kernel<<<n_segments, groups * n_bins, sizeof(float) * groups * n_bins, st>>>(feat_offset_dev, n_bins, out);
        WDR_LAUNCH_CHECK();
    }
    return WDR_OK;
}

}  // namespace wdr

using namespace wdr;

extern "C" int wdr_fbank_frames(int n_samples) { return n_samples < FB_FLEN ? 0 : 1 + (n_samples - FB_FLEN) / FB_SHIFT; }

extern "C" int wdr_kaldi_fbank_i16(const int16_t* pcm, int n, int n_bins, int subtract_mean, float* out) {
    clear_error();
    WDR_REQUIRE(n >= 0 && n_bins > 0 && n_bins <= kFbMaxBins, "bad arguments");
    const int T = wdr_fbank_frames(n);
    if (T == 0) { set_error("segment of %d samples is shorter than one 25 ms frame", n); return WDR_ERR_TOO_SHORT; }
    WDR_REQUIRE(pcm && out, "null pointer");
    int rc = ensure_device(-1);
    if (rc != WDR_OK) return rc;
    DevBuf<int16_t> d_in;
    DevBuf<float> d_out;
    DevBuf<int64_t> d_off;
    WDR_CUDA_TRY(d_in.alloc(n));
    WDR_CUDA_TRY(d_out.alloc((size_t)T * n_bins));
    WDR_CUDA_TRY(d_off.alloc(4));
    const int64_t offs[4] = {0, n, 0, T};
    WDR_CUDA_TRY(cudaMemcpy(d_in.p, pcm, sizeof(int16_t) * (size_t)n, cudaMemcpyHostToDevice));
    WDR_CUDA_TRY(cudaMemcpy(d_off.p, offs, sizeof(offs), cudaMemcpyHostToDevice));
    std::vector<int64_t> so = {0, n};
    rc = fbank_run(d_in.p, d_off.p, d_off.p + 2, so, n_bins, subtract_mean, d_out.p, 0, nullptr, nullptr);
    if (rc != WDR_OK) return rc;
    WDR_CUDA_TRY(cudaMemcpy(out, d_out.p, sizeof(float) * (size_t)T * n_bins, cudaMemcpyDeviceToHost));
    return T;
}

extern "C" int wdr_kaldi_fbank_batch_i16_dev(const int16_t* pcm, const int64_t* seg_offset, const int64_t* feat_offset,
                                             int n_segments, int64_t total_frames, int n_bins, int subtract_mean, float* out,
                                             void* stream) {
    clear_error();
    WDR_REQUIRE(n_segments >= 0 && n_bins > 0 && n_bins <= kFbMaxBins && total_frames >= 0, "bad arguments");
    if (n_segments == 0) return WDR_OK;
    WDR_REQUIRE(pcm && seg_offset && feat_offset && out, "null pointer");
    int rc = ensure_device(-1);
    if (rc != WDR_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    std::vector<int64_t> so(n_segments + 1);
    WDR_CUDA_TRY(cudaMemcpyAsync(so.data(), seg_offset, sizeof(int64_t) * (n_segments + 1), cudaMemcpyDeviceToHost, st));
    WDR_CUDA_TRY(cudaStreamSynchronize(st));
    return fbank_run(pcm, seg_offset, feat_offset, so, n_bins, subtract_mean, out, st, nullptr, nullptr);
}

extern "C" int wdr_signal_energy(const float* pcm, int n, int half_window, float* out) {
    clear_error();
    WDR_REQUIRE(n >= 0 && half_window >= 0, "bad arguments");
    if (n == 0) return WDR_OK;
    WDR_REQUIRE(pcm && out, "null pointer");
    int rc = ensure_device(-1);
    if (rc != WDR_OK) return rc;
    DevBuf<float> d_in, d_out;
    WDR_CUDA_TRY(d_in.alloc(n));
    WDR_CUDA_TRY(d_out.alloc(n));
    WDR_CUDA_TRY(cudaMemcpy(d_in.p, pcm, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice));
    int blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    signal_energy_kernel<<<blocks, 256>>>(d_in.p, n, half_window, d_out.p);
    WDR_LAUNCH_CHECK();
    WDR_CUDA_TRY(cudaMemcpy(out, d_out.p, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost));
    return WDR_OK;
}
