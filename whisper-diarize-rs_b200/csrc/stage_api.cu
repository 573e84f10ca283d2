// stage_api.cu — stage-level C entry points that are not part of whisper.h / pyannote-rs but let the parity
// tests and bench.py drive single kernels of the path (SURVEY §8b "stage-level entry points").
#include "common.cuh"
#include "gemm.cuh"
#include "ggml_file.cuh"
#include "onnx_file.cuh"
#include <string.h>

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


// ---- model-file readers without a device: what a file holds and what the loaders take from it ----
static int onnx_extract(const char* path, int kind, NamedTensors* nt, int* dim, int32_t* info) {
    OnnxFile of;
    std::string err;
    if (!path || !of.load(path, &err)) { set_error("%s", path ? err.c_str() : "null path"); return WDR_ERR_INVALID; }
    if (info) { info[0] = (int32_t)of.nodes.size(); info[1] = (int32_t)of.tensors.size(); info[2] = (int32_t)of.inputs.size(); info[3] = (int32_t)of.outputs.size(); }
    *dim = 0;
    const bool ok = kind == 0 ? onnx_extract_pyannet(of, nt, &err) : onnx_extract_resnet34(of, nt, dim, &err);
    if (!ok) { set_error("%s: %s", path, err.c_str()); return WDR_ERR_INVALID; }
    return WDR_OK;
}

extern "C" int wdr_onnx_probe(const char* path, int kind, int32_t* info) {
    clear_error();
    WDR_REQUIRE(kind == 0 || kind == 1, "kind: 0 = segmentation-3.0 (PyanNet), 1 = WeSpeaker ResNet34");
    NamedTensors nt;
    int dim = 0;
    const int rc = onnx_extract(path, kind, &nt, &dim, info);
    if (rc != WDR_OK) return rc;
    if (info) { info[4] = (int32_t)nt.size(); info[5] = dim; }
    return WDR_OK;
}

extern "C" int64_t wdr_onnx_read_param(const char* path, int kind, const char* name, float* out, int64_t cap) {
    clear_error();
    if (!name || (kind != 0 && kind != 1)) { set_error("bad arguments"); return WDR_ERR_INVALID; }
    NamedTensors nt;
    int dim = 0;
    const int rc = onnx_extract(path, kind, &nt, &dim, nullptr);
    if (rc != WDR_OK) return rc;
    auto it = nt.find(name);
    if (it == nt.end()) { set_error("no parameter '%s'", name); return WDR_ERR_INVALID; }
    const int64_t n = (int64_t)it->second.size();
    if (out && cap >= n) memcpy(out, it->second.data(), sizeof(float) * (size_t)n);
    return n;
}

extern "C" int wdr_silero_probe(const char* path, int32_t* hparams, int32_t* n_tensors) {
    clear_error();
    WDR_REQUIRE(path && hparams, "bad arguments");
    GgmlFile gf;
    SileroHeader h;
    std::string err;
    if (!gf.open_silero(path, &h, &err)) { set_error("%s", err.c_str()); return WDR_ERR_INVALID; }
    hparams[0] = h.version[0]; hparams[1] = h.version[1]; hparams[2] = h.version[2]; hparams[3] = h.n_encoder_layers;
    for (int i = 0; i < 4; i++) {
        const bool in = i < h.n_encoder_layers;
        hparams[4 + 3 * i] = in ? h.enc_in[i] : 0; hparams[5 + 3 * i] = in ? h.enc_out[i] : 0; hparams[6 + 3 * i] = in ? h.enc_kernel[i] : 0;
    }
    hparams[16] = h.lstm_input; hparams[17] = h.lstm_hidden; hparams[18] = h.final_in; hparams[19] = h.final_out;
    if (n_tensors) *n_tensors = (int32_t)gf.tensors.size();
    return WDR_OK;
}
