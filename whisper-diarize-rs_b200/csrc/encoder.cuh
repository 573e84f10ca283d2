// encoder.cuh — encoder workspace and entry points shared with the state / decoder code.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "model.cuh"
#include "profile.cuh"

namespace wdr {

struct EncoderWorkspace {
    int cap = 0;  // windows the buffers are sized for
    int64_t ldt = 0;
    float* mel_raw = nullptr;
    float* chunk_max = nullptr;
    __nv_bfloat16* frames = nullptr;
    __nv_bfloat16* conv1 = nullptr;
    float* x = nullptr;
    __nv_bfloat16* h = nullptr;
    __nv_bfloat16* qk = nullptr;
    __nv_bfloat16* vt = nullptr;
    __nv_bfloat16* att = nullptr;
    __nv_bfloat16* ff = nullptr;
    int reserve(const WhisperArch& a, int B);
    void release();
};

int encoder_forward(const wdr_context* ctx, EncoderWorkspace& ws, const float* mel, int n_frames, int mel_offset, const float* chunk_max,
                    int normalized_input, int B, float* out_f32, __nv_bfloat16* out_bf16, cudaStream_t st, Profiler* prof = nullptr);
template <typename In>
int encode_chunks(const wdr_context* ctx, EncoderWorkspace& ws, const In* pcm, int64_t chunk_stride, const int32_t* n_valid_dev, int B,
                  float* out_f32, __nv_bfloat16* out_bf16, cudaStream_t st, Profiler* prof = nullptr);
template <typename Out>
int layernorm(const float* x, const float* g, const float* b, int64_t rows, int d, Out* out, cudaStream_t st);

}  // namespace wdr
