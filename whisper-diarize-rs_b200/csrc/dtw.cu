// dtw.cu — DTW word alignment on the device.
//
// Replaces whisper.cpp `whisper_exp_compute_token_level_timestamps_dtw` pieces (`median_filter`,
// `dtw_and_backtrace`), enabled by create_context (reference src/transcribe.rs:115-136) and
// consumed through token_data().t_dtw (src/transcribe.rs:272-282).
//
//  * dtw_norm_stats / dtw_cost: ggml_norm over the token axis (double accumulation in token order, exactly
//    as ggml's CPU op), width-w median with reflect indexing, mean over heads, negate, drop sot / eot rows.
//  * dtw_wavefront: one CTA per window, one thread per text row; anti-diagonal sweep with the previous two
//    diagonals in shared memory; 2-bit trace codes packed 16 diagonals per word in a diagonal-major layout
//    (shared memory when it fits, global otherwise); serial integer backtrace by one thread, then a
//    parallel reversal.  Each cell is one fp32 add and the same three-way strict-less tie-break as the
//    sequential reference, so cost, trace and path are bit-identical to it.
#include <math.h>
#include <vector>
#include "common.cuh"

namespace wdr {

constexpr int kMaxMedianWidth = 31;

__device__ __forceinline__ int reflect_idx(int idx, int M) {
    if (idx < 0) return -idx;
    if (idx >= M) return 2 * (M - 1) - idx;
    return idx;
}

// insertion sort of a small register/local array; returns element `mid`
template <int W>
__device__ __forceinline__ float median_fixed(float* v) {
#pragma unroll
    for (int i = 1; i < W; i++) {
        float key = v[i];
        int j = i - 1;
#pragma unroll
        for (int s = 0; s < W; s++) {
            if (j >= 0 && v[j] > key) { v[j + 1] = v[j]; j--; }
        }
        v[j + 1] = key;
    }
    return v[W / 2];
}

__device__ __forceinline__ float median_any(float* v, int width) {
    for (int i = 1; i < width; i++) {
        float key = v[i];
        int j = i - 1;
        while (j >= 0 && v[j] > key) { v[j + 1] = v[j]; j--; }
        v[j + 1] = key;
    }
    return v[width / 2];
}

__global__ void median_filter_kernel(const float* __restrict__ w, int64_t rows, int M, int width, float* __restrict__ out) {
    const int64_t total = rows * M;
    const int hw = width / 2;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / M;
        const int j = (int)(e - r * M);
        const float* src = w + r * M;
        float v[kMaxMedianWidth];
        float med;
        if (width == 7) {
#pragma unroll
            for (int k = 0; k < 7; k++) v[k] = __ldg(&src[reflect_idx(j + k - 3, M)]);
            med = median_fixed<7>(v);
        } else {
            for (int k = 0; k < width; k++) v[k] = __ldg(&src[reflect_idx(j + k - hw, M)]);
            med = median_any(v, width);
        }
        out[e] = med;
    }
}

// ggml_norm statistics over the token axis for every (head, audio position): mean (float) and
// 1/sqrt(var + eps) (float), accumulated in double in token order.
__global__ void dtw_norm_stats_kernel(const float* __restrict__ w, int H, int T, int A, float* __restrict__ mean_out,
                                      float* __restrict__ scale_out) {
    const int64_t total = (int64_t)H * A;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int h = (int)(e / A), a = (int)(e % A);
        const float* col = w + (int64_t)h * T * A + a;
        double sum = 0.0;
        for (int t = 0; t < T; t++) sum += (double)col[(int64_t)t * A];
        const float mean = (float)(sum / T);
        double sum2 = 0.0;
        for (int t = 0; t < T; t++) {
            const float v = col[(int64_t)t * A] - mean;
            sum2 += (double)(v * v);
        }
        const float variance = (float)(sum2 / T);
        mean_out[e] = mean;
        scale_out[e] = 1.0f / sqrtf(variance + 1e-9f);
    }
}

__global__ void dtw_cost_kernel(const float* __restrict__ w, const float* __restrict__ mean, const float* __restrict__ scale,
                                int H, int T, int A, int sot_len, int width, float* __restrict__ out) {
    const int Nrows = T - sot_len - 1;
    const int64_t total = (int64_t)Nrows * A;
    const int hw = width / 2;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int t = (int)(e / A) + sot_len, a = (int)(e % A);
        float s = 0.0f;
        for (int h = 0; h < H; h++) {
            const float* row = w + ((int64_t)h * T + t) * A;
            const float* mh = mean + (int64_t)h * A;
            const float* sh = scale + (int64_t)h * A;
            float v[kMaxMedianWidth];
            for (int k = 0; k < width; k++) {
                const int idx = reflect_idx(a + k - hw, A);
                const float c = __ldg(&row[idx]) - __ldg(&mh[idx]);
                v[k] = c * __ldg(&sh[idx]);
            }
            s += (width == 7) ? median_fixed<7>(v) : median_any(v, width);
        }
        out[e] = -(s / (float)H);
    }
}

// ---------------------------------------------------------------------------------------------------
// wavefront DTW
// ---------------------------------------------------------------------------------------------------
struct DtwWindow {
    int64_t x_off;   // offset of x[N][M] (floats)
    int32_t N, M;
    int64_t tr_off;  // -1: trace words live in shared memory; else offset (words) into the global scratch
};

// trace words: word (q, i) holds codes of diagonals d = 16q .. 16q+15 for row i, 2 bits each.
__device__ __forceinline__ int trace_get(const uint32_t* tr, int rows, int i, int d) {
    return (tr[(size_t)(d >> 4) * rows + i] >> ((d & 15) * 2)) & 3;
}

__global__ void __launch_bounds__(1024, 1)
dtw_wavefront_kernel(const float* __restrict__ xbase, const DtwWindow* __restrict__ wins, uint32_t* __restrict__ tr_global,
                     int32_t* __restrict__ text_idx, int32_t* __restrict__ time_idx,
                     int32_t* __restrict__ path_len, int max_path, float* __restrict__ cost_out, int32_t* __restrict__ trace_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const DtwWindow win = wins[blockIdx.x];
    const int N = win.N, M = win.M;
    int32_t* ti = text_idx + (int64_t)blockIdx.x * max_path;
    int32_t* tj = time_idx + (int64_t)blockIdx.x * max_path;
    if (N <= 0 || M <= 0) {
        if (threadIdx.x == 0) path_len[blockIdx.x] = 0;
        return;
    }
    const float* x = xbase + win.x_off;
    const int rows = N + 1;
    // three rotating diagonals of N+1 floats, then (optionally) the trace words
    float* diag = reinterpret_cast<float*>(smem_raw);
    uint32_t* tr = (win.tr_off < 0) ? reinterpret_cast<uint32_t*>(diag + 3 * rows) : tr_global + win.tr_off;
    __shared__ int s_len;

    const int nthr = blockDim.x;
    // diagonal 0: cost(0,0) = 0; diagonal 1: cost(0,1) = cost(1,0) = inf
    for (int i = threadIdx.x; i < rows; i += nthr) {
        diag[0 * rows + i] = (i == 0) ? 0.0f : INFINITY;  // d = 0 (only i = 0 is a real cell)
        diag[1 * rows + i] = INFINITY;                     // d = 1
    }
    __syncthreads();

    // each thread owns rows i = tid+1, tid+1+nthr, ...; the packed trace word of the current 16-diagonal block
    // lives in registers (up to kRowsPerThread rows per thread; larger N falls back to read-modify-write).
    constexpr int kRowsPerThread = 4;
    uint32_t acc[kRowsPerThread];
#pragma unroll
    for (int r = 0; r < kRowsPerThread; r++) acc[r] = 0u;
    const bool regs_ok = (N <= kRowsPerThread * nthr);

    for (int d = 2; d <= N + M; d++) {
        const float* p2 = diag + ((d - 2) % 3) * rows;
        const float* p1 = diag + ((d - 1) % 3) * rows;
        float* cur = diag + (d % 3) * rows;
        const int i_lo = max(1, d - M), i_hi = min(N, d - 1);
        if (threadIdx.x == 0) cur[0] = INFINITY;  // cost(0, d) = inf
        auto cell = [&](int i) -> uint32_t {
            uint32_t code = 0;
            if (i >= i_lo && i <= i_hi) {
                const int j = d - i;
                const float c0 = p2[i - 1], c1 = p1[i - 1], c2 = p1[i];
                float c;
                if (c0 < c1 && c0 < c2) { c = c0; code = 0; }
                else if (c1 < c0 && c1 < c2) { c = c1; code = 1; }
                else { c = c2; code = 2; }
                const float v = __ldg(&x[(int64_t)(i - 1) * M + (j - 1)]) + c;
                cur[i] = v;
                if (cost_out) {
                    cost_out[(int64_t)i * (M + 1) + j] = v;
                    trace_out[(int64_t)i * (M + 1) + j] = (int32_t)code;
                }
            } else {
                cur[i] = INFINITY;  // off the matrix on this diagonal (j <= 0 or j > M): the inf border
            }
            return code;
        };
        const bool flush = ((d & 15) == 15) || (d == N + M);
        if (regs_ok) {
#pragma unroll
            for (int r = 0; r < kRowsPerThread; r++) {
                const int i = (int)threadIdx.x + 1 + r * nthr;
                if (i <= N) {
                    acc[r] |= cell(i) << ((d & 15) * 2);
                    if (flush) {
                        tr[(size_t)(d >> 4) * rows + i] = acc[r];
                        acc[r] = 0u;
                    }
                }
            }
        } else {
            for (int i = threadIdx.x + 1; i <= N; i += nthr) {
                const uint32_t code = cell(i);
                uint32_t* wp = &tr[(size_t)(d >> 4) * rows + i];
                const uint32_t old = ((d & 15) == 0 || d == 2) ? 0u : *wp;
                *wp = old | (code << ((d & 15) * 2));
            }
        }
        __syncthreads();
    }

    if (cost_out) {
        // borders of the full matrices (cells the sweep does not visit)
        for (int j = threadIdx.x; j <= M; j += nthr) {
            cost_out[j] = (j == 0) ? 0.0f : INFINITY;
            trace_out[j] = 2;
        }
        for (int i = threadIdx.x; i <= N; i += nthr) {
            if (i > 0) cost_out[(int64_t)i * (M + 1)] = INFINITY;
            trace_out[(int64_t)i * (M + 1)] = 1;
        }
    }
    __syncthreads();

    // serial integer backtrace (whisper.cpp: trace[0][:] = 2, trace[:][0] = 1), written reversed into the
    // tail of the output arrays, then reversed in parallel.
    if (threadIdx.x == 0) {
        int i = N, j = M, n = 0;
        while ((i > 0 || j > 0) && n < max_path) {
            ti[max_path - 1 - n] = i - 1;
            tj[max_path - 1 - n] = j - 1;
            n++;
            int t;
            if (i == 0) t = 2;
            else if (j == 0) t = 1;
            else t = trace_get(tr, rows, i, i + j);
            if (t == 0) { i--; j--; }
            else if (t == 1) { i--; }
            else { j--; }
        }
        s_len = (i > 0 || j > 0) ? -1 : n;
    }
    __syncthreads();
    const int L = s_len;
    if (L < 0) {
        if (threadIdx.x == 0) path_len[blockIdx.x] = -1;
        return;
    }
    // entries sit at [max_path - L, max_path) already in forward order (we filled from the back); move to the front
    const int shift = max_path - L;
    if (shift > 0) {
        for (int base = 0; base < L; base += nthr) {
            const int k = base + threadIdx.x;
            int a = 0, b = 0;
            if (k < L) { a = ti[shift + k]; b = tj[shift + k]; }
            __syncthreads();
            if (k < L) { ti[k] = a; tj[k] = b; }
            __syncthreads();
        }
    }
    if (threadIdx.x == 0) path_len[blockIdx.x] = L;
}

static int g_dtw_smem_by_dev[64];  // 0 = not yet queried on that device (the opt-in limit and the attribute are per device)

// Runs DTW over the windows.  x / outputs are device pointers; `wins` (host) describes the layout.
// A window keeps its packed trace in shared memory when diagonals + trace words fit (tr_off = -1),
// otherwise in a global scratch at tr_off.
int dtw_run(const float* x, std::vector<DtwWindow>& wins, int32_t* text_idx, int32_t* time_idx, int32_t* path_len,
            int max_path, float* cost_out, int32_t* trace_out, cudaStream_t st, void* wins_scratch_dev) {
    const int n = (int)wins.size();
    if (n == 0) return WDR_OK;
    int dev_cur = 0;
    WDR_CUDA_TRY(cudaGetDevice(&dev_cur));
    int& g_dtw_smem_max = g_dtw_smem_by_dev[dev_cur & 63];
    if (g_dtw_smem_max <= 0) {
        int dev = dev_cur, v = 0;
        WDR_CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        cudaFuncAttributes fa;
        WDR_CUDA_TRY(cudaFuncGetAttributes(&fa, dtw_wavefront_kernel));
        v -= (int)fa.sharedSizeBytes;  // the opt-in limit covers static + dynamic shared memory
        WDR_CUDA_TRY(cudaFuncSetAttribute(dtw_wavefront_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, v));
        g_dtw_smem_max = v;
    }
    int maxN = 0;
    size_t launch_smem = 0, global_words = 0;
    for (auto& w : wins) {
        if (w.N <= 0 || w.M <= 0) continue;
        maxN = w.N > maxN ? w.N : maxN;
        const size_t rows = (size_t)w.N + 1;
        const size_t words = ((size_t)(w.N + w.M) / 16 + 1) * rows;
        const size_t diag_bytes = 3 * rows * sizeof(float);
        const size_t full = diag_bytes + words * sizeof(uint32_t);
        if (diag_bytes > (size_t)g_dtw_smem_max) {
            set_error("dtw: N=%d exceeds the shared-memory diagonal buffers", w.N);
            return WDR_ERR_UNSUPPORTED;
        }
        size_t need;
        if (full <= (size_t)g_dtw_smem_max) {
            w.tr_off = -1;
            need = full;
        } else {
            w.tr_off = (int64_t)global_words;
            global_words += words;
            need = diag_bytes;
        }
        launch_smem = need > launch_smem ? need : launch_smem;
    }
    DtwWindow* d_wins = nullptr;
    uint32_t* d_tr = nullptr;
    if (wins_scratch_dev) d_wins = reinterpret_cast<DtwWindow*>(wins_scratch_dev);  // caller-owned, >= n entries
    else WDR_CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&d_wins), sizeof(DtwWindow) * n, st));
    WDR_CUDA_TRY(cudaMemcpyAsync(d_wins, wins.data(), sizeof(DtwWindow) * n, cudaMemcpyHostToDevice, st));
    if (global_words) WDR_CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&d_tr), sizeof(uint32_t) * global_words, st));
    int threads = ((maxN + 31) / 32) * 32;
    if (threads < 32) threads = 32;
    if (threads > 1024) threads = 1024;
    dtw_wavefront_kernel<<<n, threads, launch_smem, st>>>(x, d_wins, d_tr, text_idx, time_idx, path_len, max_path, cost_out,
                                                          trace_out);
    WDR_LAUNCH_CHECK();
    // wins.data() was read by an async copy from pageable memory: the runtime stages it before returning.
    if (!wins_scratch_dev) WDR_CUDA_TRY(cudaFreeAsync(d_wins, st));
    if (d_tr) WDR_CUDA_TRY(cudaFreeAsync(d_tr, st));
    return WDR_OK;
}

}  // namespace wdr

using namespace wdr;

extern "C" int wdr_median_filter(const float* w, int H, int N, int M, int width, float* out) {
    clear_error();
    WDR_REQUIRE(w && out && H > 0 && N > 0 && M > 0, "bad arguments");
    WDR_REQUIRE(width > 0 && (width & 1) && width <= kMaxMedianWidth, "width must be odd and <= 31");
    WDR_REQUIRE(M > width / 2, "row shorter than the half window (reflect index would leave the row)");
    int rc = ensure_device(-1);
    if (rc != WDR_OK) return rc;
    const size_t total = (size_t)H * N * M;
    DevBuf<float> d_in, d_out;
    WDR_CUDA_TRY(d_in.alloc(total));
    WDR_CUDA_TRY(d_out.alloc(total));
    WDR_CUDA_TRY(cudaMemcpy(d_in.p, w, total * sizeof(float), cudaMemcpyHostToDevice));
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    median_filter_kernel<<<blocks, 256>>>(d_in.p, (int64_t)H * N, M, width, d_out.p);
    WDR_LAUNCH_CHECK();
    WDR_CUDA_TRY(cudaMemcpy(out, d_out.p, total * sizeof(float), cudaMemcpyDeviceToHost));
    return WDR_OK;
}

namespace wdr {
int dtw_cost_dev(const float* w, int H, int T, int A, int sot_len, int width, float* mean, float* scale, float* out,
                 cudaStream_t st) {
    int blocks = (H * A + 127) / 128;
    dtw_norm_stats_kernel<<<blocks, 128, 0, st>>>(w, H, T, A, mean, scale);
    WDR_LAUNCH_CHECK();
    const int64_t total = (int64_t)(T - sot_len - 1) * A;
    blocks = (int)((total + 127) / 128);
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks > 0) {
        dtw_cost_kernel<<<blocks, 128, 0, st>>>(w, mean, scale, H, T, A, sot_len, width, out);
        WDR_LAUNCH_CHECK();
    }
    return WDR_OK;
}
}  // namespace wdr

extern "C" int wdr_dtw_cost(const float* w, int H, int n_tokens, int n_audio, int sot_len, int width, float* out) {
    clear_error();
    WDR_REQUIRE(w && out && H > 0 && n_tokens > 0 && n_audio > 0 && sot_len >= 0, "bad arguments");
    WDR_REQUIRE(n_tokens - sot_len - 1 >= 0, "n_tokens must cover sot_len + eot");
    WDR_REQUIRE(width > 0 && (width & 1) && width <= kMaxMedianWidth, "width must be odd and <= 31");
    WDR_REQUIRE(n_audio > width / 2, "n_audio shorter than the half window");
    int rc = ensure_device(-1);
    if (rc != WDR_OK) return rc;
    const size_t total = (size_t)H * n_tokens * n_audio;
    const size_t n_out = (size_t)(n_tokens - sot_len - 1) * n_audio;
    DevBuf<float> d_in, d_mean, d_scale, d_out;
    WDR_CUDA_TRY(d_in.alloc(total));
    WDR_CUDA_TRY(d_mean.alloc((size_t)H * n_audio));
    WDR_CUDA_TRY(d_scale.alloc((size_t)H * n_audio));
    WDR_CUDA_TRY(d_out.alloc(n_out));
    WDR_CUDA_TRY(cudaMemcpy(d_in.p, w, total * sizeof(float), cudaMemcpyHostToDevice));
    rc = dtw_cost_dev(d_in.p, H, n_tokens, n_audio, sot_len, width, d_mean.p, d_scale.p, d_out.p, 0);
    if (rc != WDR_OK) return rc;
    if (n_out) WDR_CUDA_TRY(cudaMemcpy(out, d_out.p, n_out * sizeof(float), cudaMemcpyDeviceToHost));
    return WDR_OK;
}

extern "C" int wdr_dtw(const float* x, int N, int M, int32_t* text_idx, int32_t* time_idx, int* path_len, float* cost_out,
                       int32_t* trace_out) {
    clear_error();
    WDR_REQUIRE(path_len && N >= 0 && M >= 0, "bad arguments");
    if (N == 0 || M == 0) { *path_len = 0; return WDR_OK; }
    WDR_REQUIRE(x && text_idx && time_idx, "null pointer");
    WDR_REQUIRE((cost_out == nullptr) == (trace_out == nullptr), "cost_out and trace_out go together");
    int rc = ensure_device(-1);
    if (rc != WDR_OK) return rc;
    const int max_path = N + M;
    const size_t cells = (size_t)(N + 1) * (M + 1);
    DevBuf<float> d_x, d_cost;
    DevBuf<int32_t> d_ti, d_tj, d_len, d_trace;
    WDR_CUDA_TRY(d_x.alloc((size_t)N * M));
    WDR_CUDA_TRY(d_ti.alloc(max_path));
    WDR_CUDA_TRY(d_tj.alloc(max_path));
    WDR_CUDA_TRY(d_len.alloc(1));
    if (cost_out) {
        WDR_CUDA_TRY(d_cost.alloc(cells));
        WDR_CUDA_TRY(d_trace.alloc(cells));
    }
    WDR_CUDA_TRY(cudaMemcpy(d_x.p, x, sizeof(float) * (size_t)N * M, cudaMemcpyHostToDevice));
    std::vector<DtwWindow> wins(1);
    wins[0].x_off = 0; wins[0].N = N; wins[0].M = M; wins[0].tr_off = 0;
    rc = dtw_run(d_x.p, wins, d_ti.p, d_tj.p, d_len.p, max_path, cost_out ? d_cost.p : nullptr, cost_out ? d_trace.p : nullptr, 0, nullptr);
    if (rc != WDR_OK) return rc;
    int32_t L = 0;
    WDR_CUDA_TRY(cudaMemcpy(&L, d_len.p, sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (L < 0) { set_error("dtw backtrace did not terminate"); return WDR_ERR_CUDA; }
    *path_len = L;
    WDR_CUDA_TRY(cudaMemcpy(text_idx, d_ti.p, sizeof(int32_t) * L, cudaMemcpyDeviceToHost));
    WDR_CUDA_TRY(cudaMemcpy(time_idx, d_tj.p, sizeof(int32_t) * L, cudaMemcpyDeviceToHost));
    if (cost_out) {
        WDR_CUDA_TRY(cudaMemcpy(cost_out, d_cost.p, sizeof(float) * cells, cudaMemcpyDeviceToHost));
        WDR_CUDA_TRY(cudaMemcpy(trace_out, d_trace.p, sizeof(int32_t) * cells, cudaMemcpyDeviceToHost));
    }
    return WDR_OK;
}

extern "C" int wdr_dtw_batch_dev(const float* x, const int64_t* x_offset, const int32_t* N, const int32_t* M, int n_windows,
                                 int32_t* text_idx, int32_t* time_idx, int32_t* path_len, int max_path, void* stream) {
    clear_error();
    WDR_REQUIRE(n_windows >= 0 && max_path > 0, "bad arguments");
    if (n_windows == 0) return WDR_OK;
    WDR_REQUIRE(x && x_offset && N && M && text_idx && time_idx && path_len, "null pointer");
    int rc = ensure_device(-1);
    if (rc != WDR_OK) return rc;
    std::vector<DtwWindow> wins(n_windows);
    for (int b = 0; b < n_windows; b++) {
        WDR_REQUIRE(N[b] >= 0 && M[b] >= 0 && N[b] + M[b] <= max_path, "window does not fit max_path");
        wins[b].x_off = x_offset[b]; wins[b].N = N[b]; wins[b].M = M[b]; wins[b].tr_off = 0;
    }
    return dtw_run(x, wins, text_idx, time_idx, path_len, max_path, nullptr, nullptr, (cudaStream_t)stream, nullptr);
}
