// state.cu — whisper_state equivalent: per-call workspaces and the encoder-level entry points.
//
// Replaces whisper_init_state / whisper_free_state (ctx.create_state(), reference src/transcribe.rs:335) and the
// encoder half of whisper_full_with_state (src/transcribe.rs:389).
#include "common.cuh"
#include "encoder.cuh"
#include "model.cuh"
#include "state.cuh"

using namespace wdr;

extern "C" wdr_state* wdr_init_state(wdr_context* ctx) {
    clear_error();
    if (!ctx) { set_error("wdr_init_state: null context"); return nullptr; }
    if (ensure_device(ctx->device) != WDR_OK) return nullptr;
    wdr_state* s = new wdr_state();
    s->ctx = ctx;
    if (cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess) {
        set_error("wdr_init_state: cudaStreamCreate failed");
        delete s;
        return nullptr;
    }
    return s;
}

extern "C" void wdr_free_state(wdr_state* s) {
    if (!s) return;
    cudaSetDevice(s->ctx->device);
    s->enc.release();
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

extern "C" int wdr_encode_chunks_i16_dev(wdr_context* ctx, wdr_state* st, const int16_t* pcm, int64_t chunk_stride,
                                         const int32_t* n_valid, int n_chunks, float* out_hidden, void* stream) {
    clear_error();
    WDR_REQUIRE(ctx && st && pcm && out_hidden && n_chunks > 0 && chunk_stride >= 0, "bad arguments");
    int rc = ensure_device(ctx->device);
    if (rc != WDR_OK) return rc;
    return encode_chunks<int16_t>(ctx, st->enc, pcm, chunk_stride, n_valid, n_chunks, out_hidden, nullptr, (cudaStream_t)stream);
}

extern "C" int wdr_encode_chunks_i16(wdr_context* ctx, wdr_state* st, const int16_t* pcm, int64_t chunk_stride, const int32_t* n_valid,
                                     int n_chunks, float* out_hidden) {
    clear_error();
    WDR_REQUIRE(ctx && st && pcm && out_hidden && n_chunks > 0 && chunk_stride >= WDR_CHUNK_SAMPLES, "bad arguments");
    int rc = ensure_device(ctx->device);
    if (rc != WDR_OK) return rc;
    const size_t n_in = (size_t)chunk_stride * (n_chunks - 1) + WDR_CHUNK_SAMPLES;
    const size_t n_out = (size_t)n_chunks * WDR_AUDIO_CTX * ctx->arch.d;
    DevBuf<int16_t> d_in;
    DevBuf<float> d_out;
    DevBuf<int32_t> d_nv;
    WDR_CUDA_TRY(d_in.alloc(n_in));
    WDR_CUDA_TRY(d_out.alloc(n_out));
    WDR_CUDA_TRY(cudaMemcpyAsync(d_in.p, pcm, sizeof(int16_t) * n_in, cudaMemcpyHostToDevice, st->stream));
    if (n_valid) {
        WDR_CUDA_TRY(d_nv.alloc(n_chunks));
        WDR_CUDA_TRY(cudaMemcpyAsync(d_nv.p, n_valid, sizeof(int32_t) * n_chunks, cudaMemcpyHostToDevice, st->stream));
    }
    rc = encode_chunks<int16_t>(ctx, st->enc, d_in.p, chunk_stride, n_valid ? d_nv.p : nullptr, n_chunks, d_out.p, nullptr, st->stream);
    if (rc != WDR_OK) return rc;
    WDR_CUDA_TRY(cudaMemcpyAsync(out_hidden, d_out.p, sizeof(float) * n_out, cudaMemcpyDeviceToHost, st->stream));
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
    rc = encoder_forward(ctx, st->enc, d_in.p, n_len, mel_offset, nullptr, 1, 1, d_out.p, nullptr, st->stream);
    if (rc != WDR_OK) return rc;
    WDR_CUDA_TRY(cudaMemcpyAsync(out_hidden, d_out.p, sizeof(float) * n_out, cudaMemcpyDeviceToHost, st->stream));
    WDR_CUDA_TRY(cudaStreamSynchronize(st->stream));
    return WDR_OK;
}
