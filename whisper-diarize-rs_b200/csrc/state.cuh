// state.cuh — wdr_state (whisper_state equivalent).
#pragma once
#include "encoder.cuh"
#include "model.cuh"

struct wdr_state {
    wdr_context* ctx = nullptr;
    cudaStream_t stream = nullptr;
    wdr::EncoderWorkspace enc;
};
