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
    // Implicit-GEMM 3 x 3 convolution (stride 1) over zero-padded 2-D maps stored as packed rows (embedding.cu: position (f, t) of a map
    // with pitch P = T + 1 is row (f + 1) * P + t of its block; row f = -1, row f = F and column t = T are zero).  The A operand is the
    // activation matrix itself — no im2col:
    //   conv2d = 1: C a multiple of 64, K = 9 C; k-block kb = (tap, 64-channel chunk); tap (dy, dx) reads the tile's rows shifted by dy P + dx;
    //   conv2d = 2: C = 32; A is viewed as OVERLAPPING 64-element rows (the channels of a position and of the next one), K = 3 x 128:
    //               k-block (dy, pair): pair 0 = taps dx = -1, 0; pair 1 = taps dx = +1, "+2" (the weights of the phantom tap are zero).
    //   conv2d = 3: plain GEMM whose output is such a map (stem, stride-2 and 1 x 1 convolutions over a materialised operand).
    // tile_pitch / tile_row0: per 128-row M tile, the pitch of the map that owns it and the row where that map's block starts (blocks are
    // multiples of 128 rows, so a tile never straddles two maps); conv_F: rows of a map.  The epilogue writes ZERO at every non-interior
    // position, so the output is again a zero-padded map.  EPI_BIAS_BF16 / EPI_BIAS_RELU_BF16 / EPI_BIAS_ADD_RELU_BF16 only.
    int conv2d = 0;
    const int32_t* tile_pitch = nullptr;
    const int32_t* tile_row0 = nullptr;
    int conv_F = 0;
};

int gemm_bf16(const GemmDesc& d, cudaStream_t st);
int gemm_last_bn();  // N-tile width (64 / 128 / 256) of the last gemm_bf16 call on this thread

}  // namespace wdr
