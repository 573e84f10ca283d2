// model.cuh — Whisper weights on the device and the per-state workspaces.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <utility>
#include <vector>
#include "../../include/wdr.h"

struct wdr_mel;

namespace wdr {

struct WhisperArch {
    const char* name;
    int d, n_head, n_enc_layer, n_dec_layer, n_mel, n_vocab;
    bool multilingual;
    int dtw_preset;  // index into the alignment-head table
};
const WhisperArch* find_arch(const char* name);

struct EncLayerW {
    float *ln1_g, *ln1_b, *ln2_g, *ln2_b;
    __nv_bfloat16* w_qkv;  // [3d][d]  rows: query | key | value
    float* b_qkv;          // [3d]     key part is zero (Whisper's key projection has no bias)
    __nv_bfloat16* w_o;    // [d][d]
    float* b_o;
    __nv_bfloat16* w_fc1;  // [4d][d]
    float* b_fc1;
    __nv_bfloat16* w_fc2;  // [d][4d]
    float* b_fc2;
};

struct DecLayerW {
    float *ln1_g, *ln1_b, *ln2_g, *ln2_b, *ln3_g, *ln3_b;
    __nv_bfloat16* w_qkv;   // self-attention [3d][d]
    float* b_qkv;
    __nv_bfloat16* w_o;
    float* b_o;
    __nv_bfloat16* w_cq;    // cross query [d][d]
    float* b_cq;
    __nv_bfloat16* w_ckv;   // cross key | value [2d][d]
    float* b_ckv;           // key part zero
    __nv_bfloat16* w_co;
    float* b_co;
    __nv_bfloat16* w_fc1;
    float* b_fc1;
    __nv_bfloat16* w_fc2;
    float* b_fc2;
};

constexpr int kConv1CPad = 128;  // mel channels padded to 128 per frame for the implicit-GEMM conv1

struct WhisperWeights {
    // encoder
    __nv_bfloat16* conv1_w;  // [d][3][128]   (tap-major, channels zero-padded)
    float* conv1_b;
    __nv_bfloat16* conv2_w;  // [d][3][d]     (tap-major)
    float* conv2_b;
    float* enc_pos;          // [1500][d] sinusoidal
    std::vector<EncLayerW> enc;
    float *enc_lnpost_g, *enc_lnpost_b;
    // decoder
    __nv_bfloat16* tok_emb;  // [round_up(n_vocab, 8)][d], pad rows zero (tied logits GEMM reads whole 8-row groups)
    float* dec_pos;          // [448][d]
    std::vector<DecLayerW> dec;
    float *dec_ln_g, *dec_ln_b;
};

}  // namespace wdr

struct wdr_context {
    int device = 0;
    wdr::WhisperArch arch;
    std::string arch_name;
    uint64_t seed = 0;
    wdr::WhisperWeights w;
    std::vector<void*> allocations;  // everything cudaMalloc'ed for the weights
    wdr_mel* mel = nullptr;
    std::vector<float> mel_filters;  // host copy [n_mel][201]
    std::vector<std::string> file_tokens;  // token strings of a ggml checkpoint (empty: synthetic vocabulary)
    std::vector<std::pair<int, int>> aheads;  // (layer, head) alignment heads of the DTW preset
    int dtw_enabled = 0;
    int dtw_preset = -1;
    size_t dtw_mem_size = 0;
    int flash_attn = 0;
    size_t weight_bytes = 0;
};
