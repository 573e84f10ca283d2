// state.cuh — wdr_state (whisper_state equivalent): stream, workspaces, device-resident results.
#pragma once
#include "encoder.cuh"
#include "model.cuh"
#include "profile.cuh"

struct wdr_state {
    wdr_context* ctx = nullptr;
    cudaStream_t stream = nullptr;       // compute
    cudaStream_t copy_stream = nullptr;  // H2D staging, overlapped with compute group by group
    wdr::EncoderWorkspace enc;
    wdr::Profiler prof;
    // device-resident staging / results of the last encode call
    int16_t* pcm_dev = nullptr;
    size_t pcm_cap = 0;
    int32_t* nvalid_dev = nullptr;
    int nvalid_cap = 0;
    float* enc_out = nullptr;  // [n_enc_chunks][1500][d] fp32
    size_t enc_out_cap = 0;
    int n_enc_chunks = 0;
    float* digest_dev = nullptr;
    int digest_cap = 0;
    std::vector<cudaEvent_t> copy_events;
};
