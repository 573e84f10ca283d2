// nn_common.cuh — small fp32 building blocks shared by the diarization networks (segmentation.cu, embedding.cu):
// counter-based weight synthesis on the host, a register-tiled SGEMM with fused bias / activation, and the persistent
// LSTM direction kernel (one CTA per (sequence, direction), W_hh rows held in registers).
//
// These networks replace the ONNX Runtime sessions pyannote-rs opens (reference src/engine.rs:117, src/transcribe.rs:343, 466).
// They are fp32 on CUDA cores: their outputs feed argmax / threshold decisions that the crate consumes as integers
// (speech state machine, cluster labels), so they keep fp32 activations end to end.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "common.cuh"

namespace wdr {

// ---- seeded weights (same generator as model.cu / the checker's weights.py) ----
inline uint64_t nn_splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
inline uint64_t nn_fnv1a(const char* s) {
    uint64_t h = 0xcbf29ce484222325ULL;
    for (; *s; s++) { h ^= (unsigned char)*s; h *= 0x100000001b3ULL; }
    return h;
}
inline std::vector<float> nn_synth(uint64_t seed, const std::string& name, size_t n, float offset, float scale) {
    const uint64_t key = nn_splitmix64(nn_fnv1a(name.c_str()) ^ nn_splitmix64(seed));
    std::vector<float> out(n);
    for (size_t i = 0; i < n; i++) {
        const uint64_t z = nn_splitmix64(key + i);
        const int k = (int)(z >> 40);
        const float u = (float)(k - 8388608) * (1.0f / 8388608.0f);
        volatile float prod = u * scale;  // separate multiply and add, as the checker computes it
        out[i] = offset + prod;
    }
    return out;
}

struct NnAllocs {
    std::vector<void*> ptrs;
    bool ok = true;
    float* upload(const std::vector<float>& h) {
        float* d = nullptr;
        if (cudaMalloc(&d, sizeof(float) * (h.empty() ? 1 : h.size())) != cudaSuccess) { ok = false; return nullptr; }
        if (!h.empty()) cudaMemcpy(d, h.data(), sizeof(float) * h.size(), cudaMemcpyHostToDevice);
        ptrs.push_back(d);
        return d;
    }
    void release() {
        for (void* p : ptrs) cudaFree(p);
        ptrs.clear();
    }
};

// Grow-only device scratch: reserve() the total once per call, then take() typed sub-buffers (256-byte aligned).
struct DevArena {
    char* base = nullptr;
    size_t cap = 0, off = 0;
    int reserve(size_t bytes) {
        off = 0;
        if (bytes <= cap) return WDR_OK;
        if (base) cudaFree(base);
        base = nullptr; cap = 0;
        WDR_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&base), bytes));
        cap = bytes;
        return WDR_OK;
    }
    static size_t padded(size_t bytes) { return (bytes + 255) & ~(size_t)255; }
    template <typename T>
    T* take(size_t count) {
        T* p = reinterpret_cast<T*>(base + off);
        off += padded(count * sizeof(T));
        return off <= cap ? p : nullptr;
    }
    void release() { if (base) cudaFree(base); base = nullptr; cap = off = 0; }
};

enum { NN_ACT_NONE = 0, NN_ACT_LEAKY = 1, NN_ACT_RELU = 2 };

// C[M][N] = act(A[M][K] * B[N][K]^T + bias[N]);  64 x 64 tile, 256 threads, 4 x 4 outputs per thread, K step 16.
static __global__ void __launch_bounds__(256)
sgemm_nt_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb, const float* __restrict__ bias,
                float* __restrict__ C, int ldc, int M, int N, int K, int act) {
    __shared__ float sA[16][64 + 4];
    __shared__ float sB[16][64 + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.0f;
    for (int k0 = 0; k0 < K; k0 += 16) {
        for (int i = tid; i < 64 * 16; i += 256) {
            const int r = i >> 4, c = i & 15;
            sA[c][r] = (m0 + r < M && k0 + c < K) ? A[(int64_t)(m0 + r) * lda + k0 + c] : 0.0f;
            sB[c][r] = (n0 + r < N && k0 + c < K) ? B[(int64_t)(n0 + r) * ldb + k0 + c] : 0.0f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; k++) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; i++) a[i] = sA[k][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; j++) b[j] = sB[k][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float v = acc[i][j] + (bias ? bias[n] : 0.0f);
            if (act == NN_ACT_LEAKY) v = v > 0.0f ? v : 0.01f * v;
            else if (act == NN_ACT_RELU) v = fmaxf(v, 0.0f);
            C[(int64_t)m * ldc + n] = v;
        }
    }
}

inline int sgemm_nt(const float* A, int lda, const float* B, int ldb, const float* bias, float* C, int ldc, int M, int N, int K, int act,
                    cudaStream_t st) {
    if (M <= 0 || N <= 0) return WDR_OK;
    sgemm_nt_kernel<<<dim3((N + 63) / 64, (M + 63) / 64), 256, 0, st>>>(A, lda, B, ldb, bias, C, ldc, M, N, K, act);
    WDR_LAUNCH_CHECK();
    return WDR_OK;
}

__device__ __forceinline__ float nn_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }

// One LSTM direction over one sequence per CTA (grid = (n_seq, n_dir)), hidden 128, 512 threads (thread r = gate row r,
// PyTorch order i|f|g|o).  gates_in[(seq*T + t) * ld_g + dir*512 + r] = W_ih x_t + b_ih + b_hh (precomputed by an SGEMM);
// out[(seq*T + t) * ld_o + dir*128 + j] = h_t[j].  dir 1 walks t = T-1 .. 0.
static __global__ void __launch_bounds__(512, 1)
lstm_dir_kernel(const float* __restrict__ gates_in, int ld_g, const float* __restrict__ whh /* [n_dir][512][128] */, int T,
                float* __restrict__ out, int ld_o) {
    __shared__ float h[128];
    __shared__ float gs[512];
    const int seq = blockIdx.x, dir = blockIdx.y, r = threadIdx.x;
    float wr[128];
    const float* wrow = whh + ((int64_t)dir * 512 + r) * 128;
#pragma unroll
    for (int j = 0; j < 128; j++) wr[j] = wrow[j];
    float c = 0.0f;
    if (r < 128) h[r] = 0.0f;
    __syncthreads();
    const float* g0 = gates_in + (int64_t)seq * T * ld_g + dir * 512 + r;
    float* o0 = out + (int64_t)seq * T * ld_o + dir * 128;
    int t = dir ? T - 1 : 0;
    const int dt = dir ? -1 : 1;
    float gnext = T > 0 ? g0[(int64_t)t * ld_g] : 0.0f;
    for (int s = 0; s < T; s++, t += dt) {
        const float a = gnext;
        if (s + 1 < T) gnext = g0[(int64_t)(t + dt) * ld_g];
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
#pragma unroll
        for (int j = 0; j < 128; j += 4) {
            a0 = fmaf(wr[j], h[j], a0);
            a1 = fmaf(wr[j + 1], h[j + 1], a1);
            a2 = fmaf(wr[j + 2], h[j + 2], a2);
            a3 = fmaf(wr[j + 3], h[j + 3], a3);
        }
        gs[r] = a + ((a0 + a1) + (a2 + a3));
        __syncthreads();
        if (r < 128) {
            const float ig = nn_sigmoid(gs[r]), fg = nn_sigmoid(gs[128 + r]), gg = tanhf(gs[256 + r]), og = nn_sigmoid(gs[384 + r]);
            c = fg * c + ig * gg;
            const float hn = og * tanhf(c);
            h[r] = hn;
            o0[(int64_t)t * ld_o + r] = hn;
        }
        __syncthreads();
    }
}

}  // namespace wdr
