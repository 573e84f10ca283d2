// mel.cu — Whisper log-mel spectrogram on the device.
//
// Replaces whisper.cpp `log_mel_spectrogram` (run inside `state.full`, reference src/transcribe.rs:389)
// and whisper_rs::convert_integer_to_float_audio (src/transcribe.rs:380-381, src/vad.rs:11-12).
//
// Layout: one CTA = 32 consecutive frames (5360 samples staged once in shared memory; every sample feeds
// 2.5 frames).  Two real frames are packed into one complex FFT-400 = 16 x 25: pass 1 (radix-16 over the
// stride-25 samples, fused Hann window + int16 decode) and pass 2 (radix-25) go through shared memory
// once each; pass 3 splits the packed spectrum into two power spectra stored bin-major/frame-minor so
// that pass 4 (sparse triangular mel filters, one warp per mel row, lane = frame) reads shared memory
// conflict-free and writes the mel-major output as coalesced 128 B rows.  The buffer-global max that
// whisper.cpp normalises with is an atomicMax per chunk; the (tiny) normalise pass is a second kernel,
// or is fused into the consumer (the encoder's conv stem) in the batched pipeline.
#include <math.h>
#include <string.h>
#include <vector>
#include "common.cuh"
#include "mel_core.cuh"

namespace wdr {

__constant__ cpx c_tw25[17] = {
    {1.f, -0.f},
    {0.968583167f, -0.24868989f},
    {0.876306653f, -0.481753677f},
    {0.72896862f, -0.684547126f},
    {0.535826802f, -0.844327927f},
    {0.309017003f, -0.95105654f},
    {0.0627905205f, -0.998026729f},
    {-0.187381312f, -0.982287228f},
    {-0.425779283f, -0.904827058f},
    {-0.637423992f, -0.770513237f},
    {-0.809017003f, -0.587785244f},
    {-0.92977649f, -0.368124545f},
    {-0.992114723f, -0.125333235f},
    {-0.992114723f, 0.125333235f},
    {-0.92977649f, 0.368124545f},
    {-0.809017003f, 0.587785244f},
    {-0.637423992f, 0.770513237f},
};

constexpr int kMelThreads = 256;
constexpr size_t kMelSmemBytes = MEL_TILE_SAMPLES * sizeof(float) + MEL_NFFT * sizeof(float) + MEL_NFFT * sizeof(cpx) +
                                 MEL_PAIRS_PER_CTA * MEL_ZPITCH * sizeof(cpx) + MEL_NBINS * MEL_PPITCH * sizeof(float);

__device__ __forceinline__ float decode_sample(float v) { return v; }
// int16 -> fp32 without I2F (which runs on the 4-lane XU pipe and held 10 % of this kernel's stall samples): 1.5 * 2^23 + v is
// exact in fp32 for |v| < 2^22, so the integer add into its bit pattern followed by the subtraction gives (float)v exactly
__device__ __forceinline__ float decode_sample(int16_t v) {
    return (__int_as_float(0x4B400000 + (int)v) - 12582912.0f) * (1.0f / 32768.0f);
}

// sample at signed position s of the reflect-padded / zero-extended signal
template <typename In>
__device__ __forceinline__ float padded_sample(const In* __restrict__ x, int n, int s) {
    if (s < 0) s = -s;  // reflect: padded[200 - i] = x[i], i = 1..200
    return (s < n) ? decode_sample(x[s]) : 0.0f;
}

template <typename In>
__global__ void __launch_bounds__(kMelThreads, 2)
log_mel_kernel(const In* __restrict__ pcm, int64_t chunk_stride, const int32_t* __restrict__ n_valid, int n_fixed,
               int n_frames, const float* __restrict__ hann_g, const cpx* __restrict__ tw400_g,
               const float* __restrict__ fw, const int4* __restrict__ frow /* {start, len, off, 0} */, int n_mel,
               float* __restrict__ out, int64_t out_chunk_stride, unsigned* __restrict__ max_keys) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* tile = reinterpret_cast<float*>(smem_raw);
    float* hann = tile + MEL_TILE_SAMPLES;
    cpx* tw400 = reinterpret_cast<cpx*>(hann + MEL_NFFT);
    cpx* zbuf = tw400 + MEL_NFFT;
    float* pbuf = reinterpret_cast<float*>(zbuf + MEL_PAIRS_PER_CTA * MEL_ZPITCH);
    __shared__ unsigned s_max;

    const int chunk = blockIdx.y;
    const int frame0 = blockIdx.x * MEL_FRAMES_PER_CTA;
    const int tid = threadIdx.x;
    const int n = n_valid ? n_valid[chunk] : n_fixed;
    const In* x = pcm + (int64_t)chunk * chunk_stride;
    float* o = out + (int64_t)chunk * out_chunk_stride;
    const int s0 = frame0 * MEL_HOP - MEL_NFFT / 2;  // signed position of tile[0]
    const int lane = tid & 31, warp = tid >> 5;
    const int frames_here = min(MEL_FRAMES_PER_CTA, n_frames - frame0);

    if (tid == 0) s_max = 0u;

    if (s0 >= n) {
        // whole tile lies in the zero padding: every mel value is log10(1e-10)
        const float floor_v = log10f(fmaxf(0.0f, 1e-10f));
        for (int i = tid; i < n_mel * MEL_FRAMES_PER_CTA; i += kMelThreads) {
            const int m = i / MEL_FRAMES_PER_CTA, f = i % MEL_FRAMES_PER_CTA;
            if (f < frames_here) o[(int64_t)m * n_frames + frame0 + f] = floor_v;
        }
        if (tid == 0 && max_keys) atomicMax(&max_keys[chunk], float_to_key(floor_v));
        return;
    }

    // ---- stage: samples (decode + reflect/zero pad), Hann window, twiddles -> shared memory ----
    constexpr int VEC = 16 / sizeof(In);  // 4 floats or 8 int16 per 128-bit load
    const bool aligned = ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    for (int v = tid; v < MEL_TILE_SAMPLES / VEC; v += kMelThreads) {
        const int j = v * VEC;
        const int s = s0 + j;
        if (aligned && s >= 0 && s + VEC <= n) {
            const int4 raw = __ldg(reinterpret_cast<const int4*>(x + s));
            const In* e = reinterpret_cast<const In*>(&raw);
#pragma unroll
            for (int q = 0; q < VEC; q++) tile[j + q] = decode_sample(e[q]);
        } else {
#pragma unroll
            for (int q = 0; q < VEC; q++) tile[j + q] = padded_sample(x, n, s + q);
        }
    }
    for (int i = tid; i < MEL_NFFT; i += kMelThreads) {
        hann[i] = hann_g[i];
        tw400[i] = tw400_g[i];
    }
    __syncthreads();

    // ---- pass 1: 16 pairs x 25 radix-16 butterflies ----
    for (int t = tid; t < MEL_PAIRS_PER_CTA * 25; t += kMelThreads) mel_pass1_task(tile, hann, tw400, zbuf, t / 25, t % 25);
    __syncthreads();
    // ---- pass 2: 16 pairs x 16 radix-25 butterflies (exactly one per thread) ----
    for (int t = tid; t < MEL_PAIRS_PER_CTA * 16; t += kMelThreads) mel_pass2_task(c_tw25, zbuf, t >> 4, t & 15);
    __syncthreads();
    // ---- pass 3: split + power ----
    for (int t = tid; t < MEL_PAIRS_PER_CTA * MEL_NBINS; t += kMelThreads) mel_pass3_task(zbuf, pbuf, t / MEL_NBINS, t % MEL_NBINS);
    __syncthreads();

    // ---- pass 4: sparse mel filterbank + log10; warp = mel row, lane = frame ----
    float vmax = -INFINITY;
    const bool live = lane < frames_here;
    for (int m = warp; m < n_mel; m += kMelThreads / 32) {
        const int4 r = __ldg(&frow[m]);
        const float* w = fw + r.z;
        const float* p = pbuf + r.x * MEL_PPITCH + lane;
        float acc = 0.0f;
        for (int k = 0; k < r.y; k++) acc = fmaf(__ldg(&w[k]), p[k * MEL_PPITCH], acc);
        const float v = log10f(fmaxf(acc, 1e-10f));
        if (live) {
            o[(int64_t)m * n_frames + frame0 + lane] = v;
            vmax = fmaxf(vmax, v);
        }
    }
    if (max_keys) {
        vmax = warp_max(vmax);
        if (lane == 0 && vmax > -INFINITY) atomicMax(&s_max, float_to_key(vmax));
        __syncthreads();
        if (tid == 0) atomicMax(&max_keys[chunk], s_max);
    }
}

// whisper.cpp normalisation: mmax = max - 8; x = (max(x, mmax) + 4) / 4.  Also decodes the max key.
__global__ void mel_normalize_kernel(float* __restrict__ mel, int64_t per_chunk, const unsigned* __restrict__ max_keys,
                                     float* __restrict__ out_max, int do_normalize) {
    const int chunk = blockIdx.y;
    const float mx = key_to_float(max_keys[chunk]);
    if (out_max && blockIdx.x == 0 && threadIdx.x == 0) out_max[chunk] = mx;
    if (!do_normalize) return;
    const float lo = mx - 8.0f;
    float* m = mel + (int64_t)chunk * per_chunk;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if ((per_chunk & 3) == 0 && (reinterpret_cast<uintptr_t>(m) & 15) == 0) {
        float4* m4 = reinterpret_cast<float4*>(m);
        for (; i < per_chunk / 4; i += stride) {
            float4 v = m4[i];
            v.x = (fmaxf(v.x, lo) + 4.0f) * 0.25f;
            v.y = (fmaxf(v.y, lo) + 4.0f) * 0.25f;
            v.z = (fmaxf(v.z, lo) + 4.0f) * 0.25f;
            v.w = (fmaxf(v.w, lo) + 4.0f) * 0.25f;
            m4[i] = v;
        }
    } else {
        for (; i < per_chunk; i += stride) m[i] = (fmaxf(m[i], lo) + 4.0f) * 0.25f;
    }
}

__global__ void i16_to_f32_kernel(const int16_t* __restrict__ in, float* __restrict__ out, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = (float)in[i] * (1.0f / 32768.0f);
}

}  // namespace wdr

using namespace wdr;

struct wdr_mel {
    int device = 0;
    int n_mel = 0;
    cudaStream_t stream = nullptr;
    float* d_hann = nullptr;
    cpx* d_tw400 = nullptr;
    float* d_fw = nullptr;
    int4* d_frow = nullptr;
    unsigned* d_maxkeys = nullptr;
    int maxkeys_cap = 0;
};

static int mel_reserve_keys(wdr_mel* m, int n_chunks) {
    if (n_chunks <= m->maxkeys_cap) return WDR_OK;
    if (m->d_maxkeys) cudaFree(m->d_maxkeys);
    m->d_maxkeys = nullptr;
    m->maxkeys_cap = 0;
    WDR_CUDA_TRY(cudaMalloc(&m->d_maxkeys, sizeof(unsigned) * n_chunks));
    m->maxkeys_cap = n_chunks;
    return WDR_OK;
}

extern "C" wdr_mel* wdr_mel_init(const float* filters, int n_mel, int device) {
    clear_error();
    if (!filters || n_mel <= 0 || n_mel > 512) { set_error("wdr_mel_init: bad arguments"); return nullptr; }
    if (ensure_device(device) != WDR_OK) return nullptr;
    wdr_mel* m = new wdr_mel();
    m->device = device;
    m->n_mel = n_mel;
    // tables in double, rounded once (whisper.cpp fills float tables from double arguments as well)
    std::vector<float> hann(MEL_NFFT);
    std::vector<cpx> tw(MEL_NFFT);
    for (int i = 0; i < MEL_NFFT; i++) {
        const double th = 2.0 * M_PI * i / MEL_NFFT;
        hann[i] = (float)(0.5 * (1.0 - cos(th)));
        tw[i].re = (float)cos(th);
        tw[i].im = (float)-sin(th);
    }
    // compact the (triangular) filter rows: [first non-zero, last non-zero]
    std::vector<int4> rows(n_mel);
    std::vector<float> fw;
    for (int r = 0; r < n_mel; r++) {
        const float* f = filters + (size_t)r * MEL_NBINS;
        int lo = MEL_NBINS, hi = -1;
        for (int k = 0; k < MEL_NBINS; k++)
            if (f[k] != 0.0f) { if (k < lo) lo = k; hi = k; }
        int len = hi >= lo ? hi - lo + 1 : 0;
        if (len == 0) lo = 0;
        rows[r] = make_int4(lo, len, (int)fw.size(), 0);
        for (int k = 0; k < len; k++) fw.push_back(f[lo + k]);
    }
    if (fw.empty()) fw.push_back(0.0f);
    bool ok = cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaMalloc(&m->d_hann, sizeof(float) * MEL_NFFT) == cudaSuccess &&
              cudaMalloc(&m->d_tw400, sizeof(cpx) * MEL_NFFT) == cudaSuccess &&
              cudaMalloc(&m->d_fw, sizeof(float) * fw.size()) == cudaSuccess &&
              cudaMalloc(&m->d_frow, sizeof(int4) * n_mel) == cudaSuccess &&
              cudaMemcpy(m->d_hann, hann.data(), sizeof(float) * MEL_NFFT, cudaMemcpyHostToDevice) == cudaSuccess &&
              cudaMemcpy(m->d_tw400, tw.data(), sizeof(cpx) * MEL_NFFT, cudaMemcpyHostToDevice) == cudaSuccess &&
              cudaMemcpy(m->d_fw, fw.data(), sizeof(float) * fw.size(), cudaMemcpyHostToDevice) == cudaSuccess &&
              cudaMemcpy(m->d_frow, rows.data(), sizeof(int4) * n_mel, cudaMemcpyHostToDevice) == cudaSuccess &&
              cudaFuncSetAttribute(log_mel_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMelSmemBytes) == cudaSuccess &&
              cudaFuncSetAttribute(log_mel_kernel<int16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMelSmemBytes) == cudaSuccess;
    if (!ok) {
        set_error("wdr_mel_init: %s", cudaGetErrorString(cudaGetLastError()));
        wdr_mel_free(m);
        return nullptr;
    }
    return m;
}

extern "C" void wdr_mel_free(wdr_mel* m) {
    if (!m) return;
    cudaSetDevice(m->device);
    if (m->stream) cudaStreamDestroy(m->stream);
    cudaFree(m->d_hann);
    cudaFree(m->d_tw400);
    cudaFree(m->d_fw);
    cudaFree(m->d_frow);
    cudaFree(m->d_maxkeys);
    delete m;
}

extern "C" int wdr_mel_n_len(int n_samples) { return (n_samples + WDR_CHUNK_SAMPLES) / WDR_HOP; }

namespace wdr {
// Launches mel (+ normalise) for `n_chunks` buffers of `n_frames` output frames each.  Device pointers.
template <typename In>
int mel_launch(wdr_mel* m, const In* pcm, int64_t chunk_stride, const int32_t* n_valid_dev, int n_fixed, int n_chunks,
               int n_frames, int normalize, float* out, float* out_max, cudaStream_t st) {
    if (n_chunks <= 0 || n_frames <= 0) return WDR_OK;
    int rc = mel_reserve_keys(m, n_chunks);
    if (rc != WDR_OK) return rc;
    WDR_CUDA_TRY(cudaMemsetAsync(m->d_maxkeys, 0, sizeof(unsigned) * n_chunks, st));
    dim3 grid((n_frames + MEL_FRAMES_PER_CTA - 1) / MEL_FRAMES_PER_CTA, n_chunks);
    const int64_t per_chunk = (int64_t)m->n_mel * n_frames;
    log_mel_kernel<In><<<grid, kMelThreads, kMelSmemBytes, st>>>(pcm, chunk_stride, n_valid_dev, n_fixed, n_frames, m->d_hann,
                                                                   m->d_tw400, m->d_fw, m->d_frow, m->n_mel, out, per_chunk,
                                                                   m->d_maxkeys);
    WDR_LAUNCH_CHECK();
    if (normalize || out_max) {
        int bx = (int)((per_chunk / 4 + 255) / 256);
        if (bx > 64) bx = 64;
        if (bx < 1) bx = 1;
        mel_normalize_kernel<<<dim3(bx, n_chunks), 256, 0, st>>>(out, per_chunk, m->d_maxkeys, out_max, normalize);
        WDR_LAUNCH_CHECK();
    }
    return WDR_OK;
}
template int mel_launch<float>(wdr_mel*, const float*, int64_t, const int32_t*, int, int, int, int, float*, float*, cudaStream_t);
template int mel_launch<int16_t>(wdr_mel*, const int16_t*, int64_t, const int32_t*, int, int, int, int, float*, float*, cudaStream_t);
}  // namespace wdr

template <typename In>
static int log_mel_host(wdr_mel* m, const In* pcm, int n, int normalize, float* out) {
    clear_error();
    WDR_REQUIRE(m && out && n >= 0 && (pcm || n == 0), "bad arguments");
    int rc = ensure_device(m->device);
    if (rc != WDR_OK) return rc;
    const int n_len = wdr_mel_n_len(n);
    DevBuf<In> d_in;
    DevBuf<float> d_out;
    WDR_CUDA_TRY(d_in.alloc((size_t)n));
    WDR_CUDA_TRY(d_out.alloc((size_t)m->n_mel * n_len));
    if (n) WDR_CUDA_TRY(cudaMemcpyAsync(d_in.p, pcm, sizeof(In) * (size_t)n, cudaMemcpyHostToDevice, m->stream));
    rc = mel_launch<In>(m, d_in.p, 0, nullptr, n, 1, n_len, normalize, d_out.p, nullptr, m->stream);
    if (rc != WDR_OK) return rc;
    WDR_CUDA_TRY(cudaMemcpyAsync(out, d_out.p, sizeof(float) * (size_t)m->n_mel * n_len, cudaMemcpyDeviceToHost, m->stream));
    WDR_CUDA_TRY(cudaStreamSynchronize(m->stream));
    return n_len;
}

extern "C" int wdr_log_mel_f32(wdr_mel* m, const float* pcm, int n, int normalize, float* out) {
    return log_mel_host<float>(m, pcm, n, normalize, out);
}
extern "C" int wdr_log_mel_i16(wdr_mel* m, const int16_t* pcm, int n, int normalize, float* out) {
    return log_mel_host<int16_t>(m, pcm, n, normalize, out);
}

extern "C" int wdr_log_mel_batch_f32_dev(wdr_mel* m, const float* pcm, int64_t chunk_stride, const int32_t* n_valid,
                                         int n_chunks, int normalize, float* out, float* out_max, void* stream) {
    clear_error();
    WDR_REQUIRE(m && pcm && out && n_chunks >= 0 && chunk_stride >= 0, "bad arguments");
    int rc = ensure_device(m->device);
    if (rc != WDR_OK) return rc;
    return mel_launch<float>(m, pcm, chunk_stride, n_valid, WDR_CHUNK_SAMPLES, n_chunks, WDR_CHUNK_FRAMES, normalize, out, out_max,
                             (cudaStream_t)stream);
}
extern "C" int wdr_log_mel_batch_i16_dev(wdr_mel* m, const int16_t* pcm, int64_t chunk_stride, const int32_t* n_valid,
                                         int n_chunks, int normalize, float* out, float* out_max, void* stream) {
    clear_error();
    WDR_REQUIRE(m && pcm && out && n_chunks >= 0 && chunk_stride >= 0, "bad arguments");
    int rc = ensure_device(m->device);
    if (rc != WDR_OK) return rc;
    return mel_launch<int16_t>(m, pcm, chunk_stride, n_valid, WDR_CHUNK_SAMPLES, n_chunks, WDR_CHUNK_FRAMES, normalize, out,
                               out_max, (cudaStream_t)stream);
}

extern "C" int wdr_log_mel_batch_i16(wdr_mel* m, const int16_t* pcm, int64_t chunk_stride, const int32_t* n_valid, int n_chunks,
                                     int normalize, float* out) {
    clear_error();
    WDR_REQUIRE(m && pcm && out && n_chunks >= 0 && chunk_stride >= WDR_CHUNK_SAMPLES, "bad arguments");
    int rc = ensure_device(m->device);
    if (rc != WDR_OK) return rc;
    if (n_chunks == 0) return WDR_OK;
    const size_t n_in = (size_t)chunk_stride * (n_chunks - 1) + WDR_CHUNK_SAMPLES;
    const size_t n_out = (size_t)n_chunks * m->n_mel * WDR_CHUNK_FRAMES;
    DevBuf<int16_t> d_in;
    DevBuf<float> d_out;
    DevBuf<int32_t> d_nv;
    WDR_CUDA_TRY(d_in.alloc(n_in));
    WDR_CUDA_TRY(d_out.alloc(n_out));
    WDR_CUDA_TRY(cudaMemcpyAsync(d_in.p, pcm, sizeof(int16_t) * n_in, cudaMemcpyHostToDevice, m->stream));
    if (n_valid) {
        WDR_CUDA_TRY(d_nv.alloc(n_chunks));
        WDR_CUDA_TRY(cudaMemcpyAsync(d_nv.p, n_valid, sizeof(int32_t) * n_chunks, cudaMemcpyHostToDevice, m->stream));
    }
    rc = mel_launch<int16_t>(m, d_in.p, chunk_stride, n_valid ? d_nv.p : nullptr, WDR_CHUNK_SAMPLES, n_chunks, WDR_CHUNK_FRAMES,
                             normalize, d_out.p, nullptr, m->stream);
    if (rc != WDR_OK) return rc;
    WDR_CUDA_TRY(cudaMemcpyAsync(out, d_out.p, sizeof(float) * n_out, cudaMemcpyDeviceToHost, m->stream));
    WDR_CUDA_TRY(cudaStreamSynchronize(m->stream));
    return WDR_OK;
}

extern "C" int wdr_convert_integer_to_float_audio(const int16_t* pcm, int n, float* out) {
    clear_error();
    WDR_REQUIRE(n >= 0 && (n == 0 || (pcm && out)), "bad arguments");
    int rc = ensure_device(-1);
    if (rc != WDR_OK) return rc;
    if (n == 0) return WDR_OK;
    DevBuf<int16_t> d_in;
    DevBuf<float> d_out;
    WDR_CUDA_TRY(d_in.alloc(n));
    WDR_CUDA_TRY(d_out.alloc(n));
    WDR_CUDA_TRY(cudaMemcpy(d_in.p, pcm, sizeof(int16_t) * (size_t)n, cudaMemcpyHostToDevice));
    i16_to_f32_kernel<<<296, 256>>>(d_in.p, d_out.p, n);
    WDR_LAUNCH_CHECK();
    WDR_CUDA_TRY(cudaMemcpy(out, d_out.p, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost));
    return WDR_OK;
}
