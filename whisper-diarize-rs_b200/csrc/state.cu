// state.cu — whisper_state equivalent: per-call workspaces and the encoder-level entry points.
//
// Replaces whisper_init_state / whisper_free_state (ctx.create_state(), reference src/transcribe.rs:335) and the
// encoder half of whisper_full_with_state (src/transcribe.rs:389).  As in whisper.cpp, the encoder output stays in
// the state on the device for the decoder; the host-pointer entry points stage PCM through a second stream in
// groups so the H2D copy of group g+1 overlaps the kernels of group g.
#include "common.cuh"
#include "encoder.cuh"
#include "model.cuh"
#include "state.cuh"

using namespace wdr;

namespace wdr {

constexpr int kEncodeGroup = 16;  // windows per H2D/compute pipeline group

template <typename T>
static int grow(T** p, size_t* cap, size_t need) {
    if (need <= *cap) return WDR_OK;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    WDR_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(p), need * sizeof(T)));
    *cap = need;
    return WDR_OK;
}

__global__ void hidden_digest_kernel(const float* __restrict__ h, int64_t per_chunk, float* __restrict__ out) {
    __shared__ float part[8];
    const float* x = h + (int64_t)blockIdx.x * per_chunk;
    float s = 0.0f;
    for (int64_t i = threadIdx.x; i < per_chunk; i += blockDim.x) s += fabsf(x[i]);
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.0f;
        for (int i = 0; i < 8; i++) t += part[i];
        out[blockIdx.x] = t / (float)per_chunk;
    }
}

}  // namespace wdr

extern "C" wdr_state* wdr_init_state(wdr_context* ctx) {
    clear_error();
    if (!ctx) { set_error("wdr_init_state: null context"); return nullptr; }
    if (ensure_device(ctx->device) != WDR_OK) return nullptr;
    wdr_state* s = new wdr_state();
    s->ctx = ctx;
    if (cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking) != cudaSuccess) {
        set_error("wdr_init_state: cudaStreamCreate failed");
        delete s;
        return nullptr;
    }
    return s;
}

extern "C" void wdr_free_state(wdr_state* s) {
    if (!s) return;
    for (auto ln : s->lanes) wdr_free_state(ln);
    s->lanes.clear();
    cudaSetDevice(s->ctx->device);
    cudaDeviceSynchronize();
    s->enc.release();
    s->dec.release();
    s->dtwp.release();
    s->full.release();
    cudaFree(s->pcm_dev);
    cudaFree(s->nvalid_dev);
    cudaFree(s->enc_out);
    cudaFree(s->digest_dev);
    for (auto e : s->copy_events) cudaEventDestroy(e);
    if (s->stream) cudaStreamDestroy(s->stream);
    if (s->copy_stream) cudaStreamDestroy(s->copy_stream);
    delete s;
}

extern "C" int wdr_encode_chunks_i16_dev(wdr_context* ctx, wdr_state* st, const int16_t* pcm, int64_t chunk_stride,
                                         const int32_t* n_valid, int n_chunks, float* out_hidden, void* stream) {
    clear_error();
    WDR_REQUIRE(ctx && st && pcm && out_hidden && n_chunks > 0 && chunk_stride >= 0, "bad arguments");
    int rc = ensure_device(ctx->device);
    if (rc != WDR_OK) return rc;
    return encode_chunks<int16_t>(ctx, st->enc, pcm, chunk_stride, n_valid, n_chunks, out_hidden, nullptr, (cudaStream_t)stream, &st->prof);
}

extern "C" int wdr_encode_chunks_i16(wdr_context* ctx, wdr_state* st, const int16_t* pcm, int64_t chunk_stride, const int32_t* n_valid,
                                     int n_chunks, float* out_hidden) {
    clear_error();
    WDR_REQUIRE(ctx && st && pcm && n_chunks > 0 && chunk_stride >= WDR_CHUNK_SAMPLES, "bad arguments");
    int rc = ensure_device(ctx->device);
    if (rc != WDR_OK) return rc;
    const int d = ctx->arch.d;
    const size_t per_out = (size_t)WDR_AUDIO_CTX * d;
    if ((rc = grow(&st->pcm_dev, &st->pcm_cap, (size_t)n_chunks * WDR_CHUNK_SAMPLES)) != WDR_OK) return rc;
    if ((rc = grow(&st->enc_out, &st->enc_out_cap, (size_t)n_chunks * per_out)) != WDR_OK) return rc;
    if (n_valid) {
        size_t cap = st->nvalid_cap;
        if ((rc = grow(&st->nvalid_dev, &cap, (size_t)n_chunks)) != WDR_OK) return rc;
        st->nvalid_cap = (int)cap;
        WDR_CUDA_TRY(cudaMemcpyAsync(st->nvalid_dev, n_valid, sizeof(int32_t) * n_chunks, cudaMemcpyHostToDevice, st->stream));
    }
    const int n_groups = (n_chunks + kEncodeGroup - 1) / kEncodeGroup;
    while ((int)st->copy_events.size() < n_groups) {
        cudaEvent_t e;
        WDR_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        st->copy_events.push_back(e);
    }
    // copy stream must not overwrite the staging buffer while a previous call's kernels still read it
    cudaEvent_t done;
    WDR_CUDA_TRY(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
    WDR_CUDA_TRY(cudaEventRecord(done, st->stream));
    WDR_CUDA_TRY(cudaStreamWaitEvent(st->copy_stream, done, 0));
    WDR_CUDA_TRY(cudaEventDestroy(done));
    for (int g = 0; g < n_groups; g++) {
        const int c0 = g * kEncodeGroup, nc = (n_chunks - c0) < kEncodeGroup ? (n_chunks - c0) : kEncodeGroup;
        if (chunk_stride == WDR_CHUNK_SAMPLES) {
            WDR_CUDA_TRY(cudaMemcpyAsync(st->pcm_dev + (size_t)c0 * WDR_CHUNK_SAMPLES, pcm + (size_t)c0 * chunk_stride,
                                         sizeof(int16_t) * (size_t)nc * WDR_CHUNK_SAMPLES, cudaMemcpyHostToDevice, st->copy_stream));
        } else {
            WDR_CUDA_TRY(cudaMemcpy2DAsync(st->pcm_dev + (size_t)c0 * WDR_CHUNK_SAMPLES, sizeof(int16_t) * WDR_CHUNK_SAMPLES,
                                           pcm + (size_t)c0 * chunk_stride, sizeof(int16_t) * chunk_stride,
                                           sizeof(int16_t) * WDR_CHUNK_SAMPLES, nc, cudaMemcpyHostToDevice, st->copy_stream));
        }
        WDR_CUDA_TRY(cudaEventRecord(st->copy_events[g], st->copy_stream));
    }
    for (int g = 0; g < n_groups; g++) {
        const int c0 = g * kEncodeGroup, nc = (n_chunks - c0) < kEncodeGroup ? (n_chunks - c0) : kEncodeGroup;
        WDR_CUDA_TRY(cudaStreamWaitEvent(st->stream, st->copy_events[g], 0));
        rc = encode_chunks<int16_t>(ctx, st->enc, st->pcm_dev + (size_t)c0 * WDR_CHUNK_SAMPLES, WDR_CHUNK_SAMPLES,
                                    n_valid ? st->nvalid_dev + c0 : nullptr, nc, st->enc_out + (size_t)c0 * per_out, nullptr, st->stream,
                                    &st->prof);
        if (rc != WDR_OK) return rc;
    }
    st->n_enc_chunks = n_chunks;
    if (out_hidden)
        WDR_CUDA_TRY(cudaMemcpyAsync(out_hidden, st->enc_out, sizeof(float) * (size_t)n_chunks * per_out, cudaMemcpyDeviceToHost, st->stream));
    WDR_CUDA_TRY(cudaStreamSynchronize(st->stream));
    return WDR_OK;
}

extern "C" int wdr_state_hidden_digest(wdr_state* st, float* out, int n) {
    clear_error();
    WDR_REQUIRE(st && out && n > 0 && n <= st->n_enc_chunks, "no encoder output of that size in the state");
    int rc = ensure_device(st->ctx->device);
    if (rc != WDR_OK) return rc;
    size_t cap = st->digest_cap;
    if ((rc = grow(&st->digest_dev, &cap, (size_t)n)) != WDR_OK) return rc;
    st->digest_cap = (int)cap;
    hidden_digest_kernel<<<n, 256, 0, st->stream>>>(st->enc_out, (int64_t)WDR_AUDIO_CTX * st->ctx->arch.d, st->digest_dev);
    WDR_LAUNCH_CHECK();
    WDR_CUDA_TRY(cudaMemcpyAsync(out, st->digest_dev, sizeof(float) * n, cudaMemcpyDeviceToHost, st->stream));
    WDR_CUDA_TRY(cudaStreamSynchronize(st->stream));
    return WDR_OK;
}

// whisper_encode semantics: mel[n_mel][n_len] already normalised (what whisper_pcm_to_mel leaves in the state),
// window starting at frame mel_offset, zero-extended to 3000 frames.  Host pointers.
extern "C" int wdr_encode(wdr_context* ctx, wdr_state* st, const float* mel, int n_len, int mel_offset, float* out_hidden) {
    clear_error();
    WDR_REQUIRE(ctx && st && mel && out_hidden && n_len > 0 && mel_offset >= 0, "bad arguments");
    int rc = ensure_device(ctx->device);
    if (rc != WDR_OK) return rc;
    rc = st->enc.reserve(ctx->arch, 1);
    if (rc != WDR_OK) return rc;
    const size_t n_in = (size_t)ctx->arch.n_mel * n_len;
    const size_t n_out = (size_t)WDR_AUDIO_CTX * ctx->arch.d;
    DevBuf<float> d_in, d_out;
    WDR_CUDA_TRY(d_in.alloc(n_in));
    WDR_CUDA_TRY(d_out.alloc(n_out));
    WDR_CUDA_TRY(cudaMemcpyAsync(d_in.p, mel, sizeof(float) * n_in, cudaMemcpyHostToDevice, st->stream));
    rc = encoder_forward(ctx, st->enc, d_in.p, n_len, mel_offset, nullptr, 1, 1, d_out.p, nullptr, st->stream, &st->prof);
    if (rc != WDR_OK) return rc;
    WDR_CUDA_TRY(cudaMemcpyAsync(out_hidden, d_out.p, sizeof(float) * n_out, cudaMemcpyDeviceToHost, st->stream));
    WDR_CUDA_TRY(cudaStreamSynchronize(st->stream));
    return WDR_OK;
}

extern "C" int wdr_profile_enable(wdr_state* st, int enable) {
    clear_error();
    WDR_REQUIRE(st, "null state");
    st->prof.enabled = enable != 0;
    for (auto ln : st->lanes) ln->prof.enabled = st->prof.enabled;
    return WDR_OK;
}

extern "C" int wdr_profile_collect(wdr_state* st, double* ms, int32_t* launches, int n_classes) {
    clear_error();
    WDR_REQUIRE(st && ms && launches && n_classes >= KC_COUNT, "need room for all kernel classes");
    for (int i = 0; i < n_classes; i++) { ms[i] = 0.0; launches[i] = 0; }
    st->prof.collect(ms, launches);
    for (auto ln : st->lanes) ln->prof.collect(ms, launches);  // lanes overlap: class sums can exceed the wall time of the call
    return KC_COUNT;
}

extern "C" int wdr_state_set_lanes(wdr_state* st, int n_lanes) {
    clear_error();
    WDR_REQUIRE(st && n_lanes >= 0 && n_lanes <= 8, "lanes must be 0 (default) .. 8");
    st->n_lanes = n_lanes;
    return WDR_OK;
}
