// gemm.cuh — host-side interface of the tcgen05 bf16 GEMM (gemm.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace wdr {

enum GemmEpilogue {
    EPI_BIAS_BF16 = 0,          // out_bf16 = acc + bias
    EPI_BIAS_GELU_BF16 = 1,     // out_bf16 = gelu(acc + bias)
    EPI_BIAS_RESID_F32 = 2,     // out_f32 = resid + acc + bias            (resid may alias out)
    EPI_BIAS_GELU_POS_F32 = 3,  // out_f32 = gelu(acc + bias) + pos[row_in_batch][n]
    EPI_QKV_BF16 = 4,           // n < n_split: out_bf16[row][n] = acc + bias;  n >= n_split: out_t[n - n_split][row] = acc + bias
    EPI_F32 = 5,                // out_f32 = acc + bias
    EPI_BIAS_GELU_SPLIT = 6,    // v = gelu_exact(acc + bias) stored as a (hi, lo) bf16 pair: out[.] = hi, out[. + split_stride] = lo (decoder fc1)
    EPI_BIAS_RELU_BF16 = 7,     // out_bf16 = relu(acc + bias)                                  (ResNet34 conv + folded BN + ReLU)
    EPI_BIAS_ADD_RELU_BF16 = 8, // out_bf16 = relu(acc + bias + resid_bf16[row][n])             (BasicBlock tail; resid_bf16 has the output's layout)
    EPI_HEADS_BF16 = 9,         // out_bf16 = acc + bias, stored head-major: row = g*group_rows + t, n = s*(N/2) + h*64 + c  ->
                                // out[((g*H + h)*2 + s)*group_rows + t][c]  (decoder cross K|V: each (window, head) K and V block contiguous)
};

// D[M x N] = A[M x K] * W[N x K]^T.  A rows are organised as n_batch groups of rows_per_batch rows (row r of
// batch b lives at A + b*a_batch_stride + r*a_row_stride elements); tiles never straddle a batch, so strided
// (overlapping) row views — the implicit-GEMM form of conv1d — are expressed by the strides alone.
// Output row index = b*rows_per_batch + r.
struct GemmDesc {
    const __nv_bfloat16* A = nullptr;
    int64_t a_row_stride = 0;    // elements; multiple of 8
    int64_t a_batch_stride = 0;  // elements; multiple of 8
    int rows_per_batch = 0;
    int n_batch = 1;
    const __nv_bfloat16* W = nullptr;  // [N][ldw] K-major
    int64_t ldw = 0;
    int N = 0, K = 0;
    // Implicit-GEMM (conv1d) addressing: the K axis is cut into "taps" of kb_per_tap 64-element blocks; tap t reads
    // columns [0, kb_per_tap*64) of A rows shifted by t (row r + t).  0 = plain GEMM (one tap spanning all of K).
    int kb_per_tap = 0;
    int a_cols = 0;  // inner extent of the A tensor (defaults to K); columns beyond it read as zero
    int epilogue = EPI_BIAS_BF16;
    void* out = nullptr;
    int64_t ldc = 0;
    int64_t c_batch_stride = 0;  // elements between batches of the output (and resid); 0 = rows_per_batch * ldc
    const float* bias = nullptr;   // [N] or null
    const float* resid = nullptr;  // fp32 [M][ldc]
    const __nv_bfloat16* resid_bf16 = nullptr;  // bf16 [M][ldc] (EPI_BIAS_ADD_RELU_BF16)
    const float* pos = nullptr;    // fp32 [rows_per_batch][N]
    __nv_bfloat16* out_t = nullptr;
    int64_t ldt = 0;
    int64_t t_batch_stride = 0;  // column distance between batches in out_t; 0 = rows_per_batch
    int n_split = 0;
    int group_rows = 0;  // EPI_HEADS_BF16: rows per group (1500 positions per window)
    // weight-streaming (decode) GEMMs: bn = 64 halves the tile width and split_k > 1 cuts K so that >= 148 CTAs stream W;
    // both require EPI_F32 without bias — split s writes its fp32 partial to out + s*split_stride (consumer sums, in order)
    int bn = 128;
    int split_k = 1;
    int64_t split_stride = 0;
    // dual_a: A is a (hi, lo) bf16 pair, lo stored a_dual_stride elements after hi; D = (hi + lo) * W^T with W read once
    bool dual_a = false;
    int64_t a_dual_stride = 0;
    // programmatic dependent launch: the kernel may start while its predecessor on the stream still runs; it prefetches its first
    // weight tiles, then waits for the predecessor before touching A / the output (decode graph only)
    bool pdl = false;
    // W is k-block-major [K/64][N][64] (decoder weights, model.cu: to_kb_major): a tile's rows are contiguous for each k-block
    bool w_kb_major = false;
};

int gemm_bf16(const GemmDesc& d, cudaStream_t st);
int gemm_last_bn();  // N-tile width (64 / 128 / 256) of the last gemm_bf16 call on this thread

}  // namespace wdr
