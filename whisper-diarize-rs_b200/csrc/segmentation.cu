// segmentation.cu — pyannote segmentation-3.0 (PyanNet) windows on the device + pyannote-rs' speech state machine on the host.
//
// Replaces pyannote_rs::get_segments(&samples, 16000, model_path) (reference src/engine.rs:117-122): 10 s windows of
// RAW int16 values cast to f32, zero-padded to a multiple of the window; per window PyanNet -> [589][7] powerset log-probs;
// per frame argmax != 0 drives a speaking / not-speaking state machine (frame_start 721, frame_size 270 samples, absolute
// across windows) that yields {start, end, samples}.
//
// Device layout for a batch of W windows (independent -> shard / batch freely):
//   x0  f32 [W][80][5325]   |sinc conv k251 s10| -> maxpool3     (InstanceNorm(1) of the waveform fused into the load)
//   x1  f32 [W][60][1773]   conv k5 -> maxpool3                  (InstanceNorm + LeakyReLU of x0 fused into the load)
//   x2  f32 [W][60][589]    conv k5 -> maxpool3
//   seq f32 [W*589][60|256] LSTM layer inputs; gates f32 [W*589][1024] (both directions from one SGEMM); 4 x biLSTM(128)
//   head: Linear 256->128 + LeakyReLU, Linear 128->128 + LeakyReLU (SGEMM epilogues), classifier 128->7 + log-softmax.
#include <math.h>
#include <string.h>
#include <vector>
#include "common.cuh"
#include "nn_common.cuh"
#include "onnx_file.cuh"

namespace wdr {

constexpr int kSegWindow = 160000, kSegFrames = 589, kSegClasses = 7;
constexpr int kT0 = 15975, kP0 = 5325, kT1 = 5321, kP1 = 1773, kT2 = 1769, kP2 = 589;
constexpr int kSegFrameStart = 721, kSegFrameSize = 270;

struct SegWeights {
    float wav_g, wav_b;
    float* conv0;            // [80][252] (k padded)
    float *n0g, *n0b;
    float *conv1, *conv1b, *n1g, *n1b;   // [60][80][5]
    float *conv2, *conv2b, *n2g, *n2b;   // [60][60][5]
    float* wih[4];           // [1024][in]   fwd rows then reverse rows
    float* bg[4];            // [1024]       b_ih + b_hh
    float* whh[4];           // [2][512][128]
    float *l0w, *l0b, *l1w, *l1b, *cw, *cb;
};

// per-window mean / rstd of the raw waveform (InstanceNorm1d(1)), double accumulation
__global__ void wav_stats_kernel(const int16_t* __restrict__ pcm, int64_t n_total, float* __restrict__ stats) {
    __shared__ double rs[32], rq[32];
    const int w = blockIdx.x;
    const int64_t base = (int64_t)w * kSegWindow;
    double s = 0.0, q = 0.0;
    for (int i = threadIdx.x; i < kSegWindow; i += blockDim.x) {
        const double v = (base + i < n_total) ? (double)pcm[base + i] : 0.0;
        s += v; q += v * v;
    }
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
    if ((threadIdx.x & 31) == 0) { rs[threadIdx.x >> 5] = s; rq[threadIdx.x >> 5] = q; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double S = 0.0, Q = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); i++) { S += rs[i]; Q += rq[i]; }
        const double mean = S / kSegWindow, var = Q / kSegWindow - mean * mean;
        stats[2 * w] = (float)mean;
        stats[2 * w + 1] = (float)(1.0 / sqrt((var > 0 ? var : 0) + 1e-5));
    }
}

// per-(window, channel) mean / rstd over T (InstanceNorm1d), double accumulation.  x: [W][C][T]
__global__ void chan_stats_kernel(const float* __restrict__ x, int T, float* __restrict__ stats) {
    __shared__ double rs[32], rq[32];
    const int64_t row = (int64_t)blockIdx.y * gridDim.x + blockIdx.x;
    const float* p = x + row * T;
    double s = 0.0, q = 0.0;
    for (int i = threadIdx.x; i < T; i += blockDim.x) { const double v = p[i]; s += v; q += v * v; }
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
    if ((threadIdx.x & 31) == 0) { rs[threadIdx.x >> 5] = s; rq[threadIdx.x >> 5] = q; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double S = 0.0, Q = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); i++) { S += rs[i]; Q += rq[i]; }
        const double mean = S / T, var = Q / T - mean * mean;
        stats[2 * row] = (float)mean;
        stats[2 * row + 1] = (float)(1.0 / sqrt((var > 0 ? var : 0) + 1e-5));
    }
}

// sinc conv (1 -> 80, k 251, stride 10) + |.| + maxpool3.  CTA = 64 pooled outputs of one window, all 80 channels.
constexpr int kSincTile = 64, kSincIn = (kSincTile * 3 - 1) * 10 + 251;  // 2161 input samples
constexpr size_t kSincSmem = sizeof(float) * (80 * 252 + kSincIn + 3);
__global__ void __launch_bounds__(256)
sinc_conv_pool_kernel(const int16_t* __restrict__ pcm, int64_t n_total, const float* __restrict__ wstats, SegWeights w, float* __restrict__ out) {
    extern __shared__ float sm[];
    float* f = sm;                // [80][252]
    float* xin = sm + 80 * 252;   // [kSincIn]
    const int win = blockIdx.y, tp0 = blockIdx.x * kSincTile, tid = threadIdx.x;
    for (int i = tid; i < 80 * 252; i += 256) f[i] = w.conv0[i];
    const float mean = wstats[2 * win], rstd = wstats[2 * win + 1];
    const int64_t base = (int64_t)win * kSegWindow + (int64_t)tp0 * 30;
    for (int i = tid; i < kSincIn; i += 256) {
        const int64_t s = base + i;
        const bool in_win = (tp0 * 30 + i) < kSegWindow;
        const float v = (in_win && s < n_total) ? (float)pcm[s] : 0.0f;
        xin[i] = in_win ? (v - mean) * rstd * w.wav_g + w.wav_b : 0.0f;
    }
    __syncthreads();
    const int tp = tid & 63, cg = tid >> 6;  // 4 channel groups
    if (tp0 + tp >= kP0) return;
    const float* x0 = xin + tp * 30;
    for (int c = cg; c < 80; c += 4) {
        const float* fc = f + c * 252;
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f;
#pragma unroll 4
        for (int k = 0; k < 251; k++) {
            const float fv = fc[k];
            a0 = fmaf(fv, x0[k], a0);
            a1 = fmaf(fv, x0[k + 10], a1);
            a2 = fmaf(fv, x0[k + 20], a2);
        }
        out[((int64_t)win * 80 + c) * kP0 + tp0 + tp] = fmaxf(fabsf(a0), fmaxf(fabsf(a1), fabsf(a2)));
    }
}

// conv1d (C_in -> C_out, k 5, stride 1) + bias + maxpool3 with InstanceNorm + LeakyReLU of the input fused into the load.
// CTA = 32 pooled outputs (lane) of one window; warps stride over output channels; the normalised input tile lives in smem.
template <int CIN>
__global__ void __launch_bounds__(256)
conv5_pool_kernel(const float* __restrict__ x, int T_in, const float* __restrict__ stats, const float* __restrict__ g, const float* __restrict__ b,
                  const float* __restrict__ wt /* [C_out][CIN][5] */, const float* __restrict__ bias, int C_out, int P_out, float* __restrict__ out) {
    __shared__ float xs[CIN][104];  // 32*3 + 4 = 100 time steps
    const int win = blockIdx.y, tp0 = blockIdx.x * 32, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < CIN * 100; i += 256) {
        const int c = i / 100, t = i % 100;
        const int ti = tp0 * 3 + t;
        float v = 0.0f;
        if (ti < T_in) {
            const int64_t row = (int64_t)win * CIN + c;
            v = (x[row * T_in + ti] - stats[2 * row]) * stats[2 * row + 1] * g[c] + b[c];
            v = v > 0.0f ? v : 0.01f * v;
        }
        xs[c][t] = v;
    }
    __syncthreads();
    if (tp0 + lane >= P_out) return;
    for (int co = warp; co < C_out; co += 8) {
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f;
        const float* wr = wt + (int64_t)co * CIN * 5;
        for (int ci = 0; ci < CIN; ci++) {
            const float* xr = &xs[ci][lane * 3];
#pragma unroll
            for (int k = 0; k < 5; k++) {
                const float wv = __ldg(&wr[ci * 5 + k]);
                a0 = fmaf(wv, xr[k], a0);
                a1 = fmaf(wv, xr[k + 1], a1);
                a2 = fmaf(wv, xr[k + 2], a2);
            }
        }
        out[((int64_t)win * C_out + co) * P_out + tp0 + lane] = fmaxf(a0, fmaxf(a1, a2)) + bias[co];
    }
}

// leaky(IN(x2)) transposed to the LSTM input layout: x2 [W][60][589] -> seq [W*589][60]
__global__ void norm_transpose_kernel(const float* __restrict__ x, const float* __restrict__ stats, const float* __restrict__ g,
                                      const float* __restrict__ b, float* __restrict__ seq) {
    const int win = blockIdx.y;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 60 * kSegFrames; i += gridDim.x * blockDim.x) {
        const int t = i / 60, c = i % 60;
        const int64_t row = (int64_t)win * 60 + c;
        float v = (x[row * kSegFrames + t] - stats[2 * row]) * stats[2 * row + 1] * g[c] + b[c];
        v = v > 0.0f ? v : 0.01f * v;
        seq[((int64_t)win * kSegFrames + t) * 60 + c] = v;
    }
}

// classifier 128 -> 7 + log-softmax: one warp per frame
__global__ void classifier_kernel(const float* __restrict__ y, const float* __restrict__ cw, const float* __restrict__ cb, int64_t n_rows,
                                  float* __restrict__ scores) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= n_rows) return;
    float z[kSegClasses];
    const float* yr = y + row * 128;
    const float y0 = yr[lane], y1 = yr[lane + 32], y2 = yr[lane + 64], y3 = yr[lane + 96];
#pragma unroll
    for (int k = 0; k < kSegClasses; k++) {
        const float* wr = cw + k * 128;
        float a = wr[lane] * y0 + wr[lane + 32] * y1 + wr[lane + 64] * y2 + wr[lane + 96] * y3;
        z[k] = warp_sum(a) + cb[k];
    }
    if (lane == 0) {
        float m = z[0];
#pragma unroll
        for (int k = 1; k < kSegClasses; k++) m = fmaxf(m, z[k]);
        float s = 0.0f;
#pragma unroll
        for (int k = 0; k < kSegClasses; k++) s += expf(z[k] - m);
        const float ls = logf(s);
#pragma unroll
        for (int k = 0; k < kSegClasses; k++) scores[row * kSegClasses + k] = z[k] - m - ls;
    }
}

}  // namespace wdr

using namespace wdr;

struct wdr_seg {
    int device = 0;
    cudaStream_t stream = nullptr;
    SegWeights w;
    NnAllocs mem;
    float* arena = nullptr;   // grow-only activation workspace of seg_forward (cudaMalloc / cudaFree of ~100 MB per call cost more
    size_t arena_cap = 0;     // than the network itself)
    int16_t* pcm_dev = nullptr;
    size_t pcm_cap = 0;
    float* scores_dev = nullptr;
    size_t scores_cap = 0;
};
struct wdr_seg_result {
    std::vector<double> start, end;
    std::vector<int64_t> i0, i1;
    std::vector<int16_t> padded;  // the zero-padded input the sample ranges index into
};

extern "C" wdr_seg* wdr_seg_init(const char* path, uint64_t seed, int device) {
    clear_error();
    // pyannote_rs::get_segments(.., model_path) (src/engine.rs:117-122): segmentation-3.0.onnx; NULL / "" = seeded weights
    NamedTensors file_w;
    const bool from_file = path && path[0];
    if (from_file) {
        OnnxFile of;
        std::string err;
        if (!of.load(path, &err) || !onnx_extract_pyannet(of, &file_w, &err)) { set_error("wdr_seg_init: %s", err.c_str()); return nullptr; }
    }
    if (ensure_device(device) != WDR_OK) return nullptr;
    wdr_seg* m = new wdr_seg();
    m->device = device;
    if (cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking) != cudaSuccess) { set_error("stream"); delete m; return nullptr; }
    SegWeights& w = m->w;
    bool file_ok = true;
    auto S = [&](const char* n, size_t cnt, float off, float sc) {
        if (!from_file) return nn_synth(seed, std::string("pyannet.") + n, cnt, off, sc);
        auto it = file_w.find(n);  // same PyTorch-style names and layouts (onnx_extract_pyannet)
        if (it == file_w.end() || it->second.size() != cnt) { file_ok = false; return std::vector<float>(cnt, 0.0f); }
        return it->second;
    };
    w.wav_g = S("wav_norm.weight", 1, 1.0f, 0.1f)[0];
    w.wav_b = S("wav_norm.bias", 1, 0.0f, 0.1f)[0];
    {
        std::vector<float> c0 = S("conv0.weight", 80 * 251, 0.0f, (float)(1.0 / sqrt(251.0))), p(80 * 252, 0.0f);
        for (int c = 0; c < 80; c++) memcpy(&p[c * 252], &c0[c * 251], sizeof(float) * 251);
        w.conv0 = m->mem.upload(p);
    }
    w.n0g = m->mem.upload(S("norm0.weight", 80, 1.0f, 0.1f));
    w.n0b = m->mem.upload(S("norm0.bias", 80, 0.0f, 0.1f));
    const float s1 = (float)(1.0 / sqrt(80.0 * 5)), s2 = (float)(1.0 / sqrt(60.0 * 5));
    w.conv1 = m->mem.upload(S("conv1.weight", 60 * 80 * 5, 0.0f, s1));
    w.conv1b = m->mem.upload(S("conv1.bias", 60, 0.0f, s1));
    w.n1g = m->mem.upload(S("norm1.weight", 60, 1.0f, 0.1f));
    w.n1b = m->mem.upload(S("norm1.bias", 60, 0.0f, 0.1f));
    w.conv2 = m->mem.upload(S("conv2.weight", 60 * 60 * 5, 0.0f, s2));
    w.conv2b = m->mem.upload(S("conv2.bias", 60, 0.0f, s2));
    w.n2g = m->mem.upload(S("norm2.weight", 60, 1.0f, 0.1f));
    w.n2b = m->mem.upload(S("norm2.bias", 60, 0.0f, 0.1f));
    // Recurrent / head matrices are drawn 2.5x / 2x / 6x wider than PyTorch's default init and the "no speaker" class gets a
    // +1.5 bias: with default-init scales a random PyanNet's output is constant in time (no segment is ever emitted); with these
    // the speech state machine flips a few times per window, so get_segments has real work (the CPU checker uses the same constants).
    const float sl = (float)(1.0 / sqrt(128.0)), slw = (float)(2.5 / sqrt(128.0));
    for (int l = 0; l < 4; l++) {
        const int n_in = l == 0 ? 60 : 256;
        std::vector<float> wih, bg, whh;
        for (const char* d : {"", "_reverse"}) {
            char nm[64];
            snprintf(nm, sizeof(nm), "lstm.weight_ih_l%d%s", l, d);
            auto a = S(nm, (size_t)512 * n_in, 0.0f, slw);
            wih.insert(wih.end(), a.begin(), a.end());
            snprintf(nm, sizeof(nm), "lstm.weight_hh_l%d%s", l, d);
            auto h = S(nm, 512 * 128, 0.0f, slw);
            whh.insert(whh.end(), h.begin(), h.end());
            snprintf(nm, sizeof(nm), "lstm.bias_ih_l%d%s", l, d);
            auto b1 = S(nm, 512, 0.0f, sl);
            snprintf(nm, sizeof(nm), "lstm.bias_hh_l%d%s", l, d);
            auto b2 = S(nm, 512, 0.0f, sl);
            for (int i = 0; i < 512; i++) bg.push_back(b1[i] + b2[i]);
        }
        w.wih[l] = m->mem.upload(wih);
        w.bg[l] = m->mem.upload(bg);
        w.whh[l] = m->mem.upload(whh);
    }
    w.l0w = m->mem.upload(S("linear0.weight", 128 * 256, 0.0f, 2.0f / 16));
    w.l0b = m->mem.upload(S("linear0.bias", 128, 0.0f, 1.0f / 16));
    w.l1w = m->mem.upload(S("linear1.weight", 128 * 128, 0.0f, (float)(2.0 / sqrt(128.0))));
    w.l1b = m->mem.upload(S("linear1.bias", 128, 0.0f, sl));
    w.cw = m->mem.upload(S("classifier.weight", 7 * 128, 0.0f, (float)(24.0 / sqrt(128.0))));
    {
        auto cb = S("classifier.bias", 7, 0.0f, 0.5f);
        if (!from_file) cb[0] += 1.5f;  // seeded weights only (see above)
        w.cb = m->mem.upload(cb);
    }
    if (!file_ok) { set_error("wdr_seg_init: the model file lacks a PyanNet parameter"); wdr_seg_free(m); return nullptr; }
    if (!m->mem.ok || cudaFuncSetAttribute(sinc_conv_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSincSmem) != cudaSuccess) {
        set_error("wdr_seg_init: device allocation failed");
        wdr_seg_free(m);
        return nullptr;
    }
    return m;
}

extern "C" void wdr_seg_free(wdr_seg* m) {
    if (!m) return;
    cudaSetDevice(m->device);
    cudaDeviceSynchronize();
    m->mem.release();
    cudaFree(m->arena);
    cudaFree(m->pcm_dev);
    cudaFree(m->scores_dev);
    if (m->stream) cudaStreamDestroy(m->stream);
    delete m;
}

// pyannote-rs pads `window_size - (len % window_size)` zeros: floor(n / window) + 1 windows — an exact multiple of 10 s gets a whole
// extra silent window (which closes a speaker still active at the end of the audio), n = 0 gets one
extern "C" int wdr_seg_n_windows(int64_t n_samples) { return n_samples < 0 ? 0 : (int)(n_samples / kSegWindow + 1); }

// scores[W][589][7] (device) for W windows of pcm (device int16, n_total valid samples; the tail of the last window reads as 0)
static int seg_forward(wdr_seg* m, const int16_t* pcm_dev, int64_t n_total, int W, float* scores_dev, cudaStream_t st) {
    const SegWeights& w = m->w;
    const int64_t R = (int64_t)W * kSegFrames;
    struct Sub { float* p; };
    const size_t sizes[12] = {(size_t)2 * W, (size_t)W * 80 * kP0, (size_t)W * 80 * 2, (size_t)W * 60 * kP1, (size_t)W * 60 * 2, (size_t)W * 60 * kP2,
                              (size_t)W * 60 * 2, (size_t)R * 256, (size_t)R * 256, (size_t)R * 1024, (size_t)R * 128, (size_t)R * 128};
    size_t total = 0;
    for (size_t n : sizes) total += (n + 63) & ~(size_t)63;
    if (total > m->arena_cap) {
        if (m->arena) cudaFree(m->arena);
        m->arena = nullptr; m->arena_cap = 0;
        WDR_CUDA_TRY(cudaMalloc(&m->arena, sizeof(float) * total));
        m->arena_cap = total;
    }
    size_t off = 0;
    int si = 0;
    auto carve = [&]() { Sub b{m->arena + off}; off += (sizes[si++] + 63) & ~(size_t)63; return b; };
    Sub wst = carve(), x0 = carve(), st0 = carve(), x1 = carve(), st1 = carve(), x2 = carve(), st2 = carve(), seqA = carve(), seqB = carve(),
        gates = carve(), y0 = carve(), y1 = carve();
    wav_stats_kernel<<<W, 1024, 0, st>>>(pcm_dev, n_total, wst.p);
    WDR_LAUNCH_CHECK();
    sinc_conv_pool_kernel<<<dim3((kP0 + kSincTile - 1) / kSincTile, W), 256, kSincSmem, st>>>(pcm_dev, n_total, wst.p, w, x0.p);
    WDR_LAUNCH_CHECK();
    chan_stats_kernel<<<dim3(80, W), 256, 0, st>>>(x0.p, kP0, st0.p);
    WDR_LAUNCH_CHECK();
    conv5_pool_kernel<80><<<dim3((kP1 + 31) / 32, W), 256, 0, st>>>(x0.p, kP0, st0.p, w.n0g, w.n0b, w.conv1, w.conv1b, 60, kP1, x1.p);
    WDR_LAUNCH_CHECK();
    chan_stats_kernel<<<dim3(60, W), 256, 0, st>>>(x1.p, kP1, st1.p);
    WDR_LAUNCH_CHECK();
    conv5_pool_kernel<60><<<dim3((kP2 + 31) / 32, W), 256, 0, st>>>(x1.p, kP1, st1.p, w.n1g, w.n1b, w.conv2, w.conv2b, 60, kP2, x2.p);
    WDR_LAUNCH_CHECK();
    chan_stats_kernel<<<dim3(60, W), 256, 0, st>>>(x2.p, kP2, st2.p);
    WDR_LAUNCH_CHECK();
    norm_transpose_kernel<<<dim3(32, W), 256, 0, st>>>(x2.p, st2.p, w.n2g, w.n2b, seqA.p);
    WDR_LAUNCH_CHECK();
    float* cur = seqA.p;
    float* nxt = seqB.p;
    int rc;
    for (int l = 0; l < 4; l++) {
        const int n_in = l == 0 ? 60 : 256;
        if ((rc = sgemm_nt(cur, n_in, w.wih[l], n_in, w.bg[l], gates.p, 1024, (int)R, 1024, n_in, NN_ACT_NONE, st)) != WDR_OK) return rc;
        lstm_dir_kernel<<<dim3(W, 2), 512, 0, st>>>(gates.p, 1024, w.whh[l], kSegFrames, nxt, 256);
        WDR_LAUNCH_CHECK();
        float* t = cur; cur = nxt; nxt = t;
    }
    if ((rc = sgemm_nt(cur, 256, w.l0w, 256, w.l0b, y0.p, 128, (int)R, 128, 256, NN_ACT_LEAKY, st)) != WDR_OK) return rc;
    if ((rc = sgemm_nt(y0.p, 128, w.l1w, 128, w.l1b, y1.p, 128, (int)R, 128, 128, NN_ACT_LEAKY, st)) != WDR_OK) return rc;
    classifier_kernel<<<(unsigned)((R + 7) / 8), 256, 0, st>>>(y1.p, w.cw, w.cb, R, scores_dev);
    WDR_LAUNCH_CHECK();
    return WDR_OK;  // asynchronous: the arena belongs to the model and the next call runs on the same stream
}

extern "C" int wdr_seg_scores_i16(wdr_seg* m, const int16_t* pcm, int64_t n, float* scores) {
    clear_error();
    WDR_REQUIRE(m && n >= 0 && (pcm || n == 0) && scores, "bad arguments");
    int rc = ensure_device(m->device);
    if (rc != WDR_OK) return rc;
    const int W = wdr_seg_n_windows(n);
    struct { int16_t* p; } d_x;
    struct { float* p; } d_s;
    if ((size_t)n > m->pcm_cap) {
        if (m->pcm_dev) cudaFree(m->pcm_dev);
        m->pcm_dev = nullptr; m->pcm_cap = 0;
        WDR_CUDA_TRY(cudaMalloc(&m->pcm_dev, sizeof(int16_t) * (size_t)n));
        m->pcm_cap = (size_t)n;
    }
    const size_t n_scores = (size_t)W * kSegFrames * kSegClasses;
    if (n_scores > m->scores_cap) {
        if (m->scores_dev) cudaFree(m->scores_dev);
        m->scores_dev = nullptr; m->scores_cap = 0;
        WDR_CUDA_TRY(cudaMalloc(&m->scores_dev, sizeof(float) * n_scores));
        m->scores_cap = n_scores;
    }
    d_x.p = m->pcm_dev;
    d_s.p = m->scores_dev;
    if (n) WDR_CUDA_TRY(cudaMemcpyAsync(d_x.p, pcm, sizeof(int16_t) * (size_t)n, cudaMemcpyHostToDevice, m->stream));
    // windows are independent: process them in groups to bound the workspace (x0 alone is 1.7 MB per window)
    const int group = 64;
    for (int w0 = 0; w0 < W; w0 += group) {
        const int nw = W - w0 < group ? W - w0 : group;
        rc = seg_forward(m, d_x.p + (int64_t)w0 * kSegWindow, n - (int64_t)w0 * kSegWindow, nw, d_s.p + (size_t)w0 * kSegFrames * kSegClasses, m->stream);
        if (rc != WDR_OK) return rc;
    }
    WDR_CUDA_TRY(cudaMemcpyAsync(scores, d_s.p, sizeof(float) * n_scores, cudaMemcpyDeviceToHost, m->stream));
    WDR_CUDA_TRY(cudaStreamSynchronize(m->stream));
    return W;
}

// pyannote-rs' state machine on [n_windows][589][7] scores (host logic, bit-exact given the scores).  n = the ORIGINAL sample count:
// upstream clamps the sample range to it (start to n - 1, end to n), not to the padded length, after a seconds round trip in f64
// (start = offset / sr; idx = (start * sr) as usize), which is restated literally because the truncation can land one below offset.
static void seg_state_machine(const float* scores, int W, int64_t n, wdr_seg_result* out) {
    int64_t offset = kSegFrameStart, start = 0;
    bool speaking = false;
    for (int64_t f = 0; f < (int64_t)W * kSegFrames; f++) {
        const float* z = scores + f * kSegClasses;
        int cls = 0;
        for (int k = 1; k < kSegClasses; k++)
            if (z[k] > z[cls]) cls = k;  // first maximum
        if (cls != 0) {
            if (!speaking) { start = offset; speaking = true; }
        } else if (speaking) {
            const double sr = (double)WDR_SAMPLE_RATE;
            const double t0 = (double)start / sr, t1 = (double)offset / sr;
            out->start.push_back(t0);
            out->end.push_back(t1);
            const double lim0 = (double)(n > 0 ? n - 1 : 0), lim1 = (double)n;
            const double s_f = t0 * sr, e_f = t1 * sr;
            int64_t a = (int64_t)(s_f < lim0 ? s_f : lim0), b = (int64_t)(e_f < lim1 ? e_f : lim1);
            if (b < a) b = a;  // upstream would panic on start > end; cannot happen for n >= 1
            out->i0.push_back(a);
            out->i1.push_back(b);
            speaking = false;
        }
        offset += kSegFrameSize;
    }
}

extern "C" wdr_seg_result* wdr_seg_segments_from_scores(const float* scores, int n_windows, int64_t n_samples) {
    clear_error();
    if (n_windows < 0 || n_samples < 0 || (!scores && n_windows)) { set_error("bad arguments"); return nullptr; }
    wdr_seg_result* r = new wdr_seg_result();
    seg_state_machine(scores, n_windows, n_samples, r);
    return r;
}

extern "C" wdr_seg_result* wdr_seg_get_segments(wdr_seg* m, const int16_t* pcm, int64_t n) {
    clear_error();
    if (!m || n < 0 || (!pcm && n)) { set_error("bad arguments"); return nullptr; }
    const int W = wdr_seg_n_windows(n);
    std::vector<float> scores((size_t)W * kSegFrames * kSegClasses);
    if (W && wdr_seg_scores_i16(m, pcm, n, scores.data()) < 0) return nullptr;
    wdr_seg_result* r = new wdr_seg_result();
    r->padded.assign((size_t)W * kSegWindow, 0);
    if (n) memcpy(r->padded.data(), pcm, sizeof(int16_t) * (size_t)n);
    seg_state_machine(scores.data(), W, n, r);
    return r;
}
extern "C" int wdr_seg_result_n(wdr_seg_result* r) { return r ? (int)r->start.size() : 0; }
extern "C" double wdr_seg_result_start(wdr_seg_result* r, int i) { return (r && i >= 0 && i < (int)r->start.size()) ? r->start[i] : -1.0; }
extern "C" double wdr_seg_result_end(wdr_seg_result* r, int i) { return (r && i >= 0 && i < (int)r->end.size()) ? r->end[i] : -1.0; }
extern "C" int64_t wdr_seg_result_sample_range(wdr_seg_result* r, int i, int64_t* i1) {
    if (!r || i < 0 || i >= (int)r->i0.size()) return -1;
    if (i1) *i1 = r->i1[i];
    return r->i0[i];
}
extern "C" const int16_t* wdr_seg_result_samples(wdr_seg_result* r, int i, int64_t* count) {
    if (!r || i < 0 || i >= (int)r->i0.size() || r->padded.empty()) { if (count) *count = 0; return nullptr; }
    if (count) *count = r->i1[i] - r->i0[i];
    return r->padded.data() + r->i0[i];
}
extern "C" void wdr_seg_result_free(wdr_seg_result* r) { delete r; }
