// stage_api.cu — stage-level C entry points that are not part of whisper.h / pyannote-rs but let the parity
// tests and bench.py drive single kernels of the path (SURVEY §8b "stage-level entry points").
#include "common.cuh"
#include "gemm.cuh"

using namespace wdr;

extern "C" int wdr_gemm_bf16_dev(const uint16_t* A, int64_t lda, int rows_per_batch, int n_batch, int64_t a_batch_stride,
                                 const uint16_t* W, int64_t ldw, int N, int K, int kb_per_tap, int a_cols, const float* bias,
                                 int epilogue, void* out, int64_t ldc, const float* resid_or_pos, uint16_t* out_t, int64_t ldt,
                                 int n_split, int64_t t_batch_stride, void* stream) {
    clear_error();
    int rc = ensure_device(-1);
    if (rc != WDR_OK) return rc;
    GemmDesc d;
    d.A = reinterpret_cast<const __nv_bfloat16*>(A);
    d.a_row_stride = lda;
    d.a_batch_stride = a_batch_stride;
    d.rows_per_batch = rows_per_batch;
    d.n_batch = n_batch;
    d.W = reinterpret_cast<const __nv_bfloat16*>(W);
    d.ldw = ldw;
    d.N = N;
    d.K = K;
    d.kb_per_tap = kb_per_tap;
    d.a_cols = a_cols;
    d.epilogue = epilogue;
    d.out = out;
    d.ldc = ldc;
    d.bias = bias;
    d.resid = (epilogue == EPI_BIAS_RESID_F32) ? resid_or_pos : nullptr;
    d.pos = (epilogue == EPI_BIAS_GELU_POS_F32) ? resid_or_pos : nullptr;
    d.out_t = reinterpret_cast<__nv_bfloat16*>(out_t);
    d.ldt = ldt;
    d.n_split = n_split;
    d.t_batch_stride = t_batch_stride;
    if (epilogue == EPI_HEADS_BF16) { d.group_rows = n_split; d.n_split = 0; }  // head-major K|V store: n_split carries the rows per group
    return gemm_bf16(d, (cudaStream_t)stream);
}

extern "C" int wdr_gemm_last_tile_n(void) { return gemm_last_bn(); }
