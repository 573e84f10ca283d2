// encoder.cu — the Whisper audio encoder, batched over 30 s windows.
//
// Replaces whisper.cpp `whisper_encode_internal` (conv stem + n_layer pre-LN transformer blocks + ln_post;
// SURVEY A.2), which `state.full` runs once per window (reference src/transcribe.rs:389).
//
// Data layout in HBM for a batch of B windows (T = 1500 positions, d = model width):
//   mel_raw  f32  [B][n_mel][3000]      raw log10 mel + per-window max (mel kernel)
//   frames   bf16 [B][3002][128]        normalised mel, frame-major, one zero row either side, channels padded
//   conv1    bf16 [B][3002][d]          GELU(conv1), row 0 / 3001 zero  (implicit GEMM, 3 taps x 128)
//   x        f32  [B*T][d]              residual stream (conv2 as implicit GEMM over row pairs, + sinusoids)
//   h        bf16 [B*T][d]              LayerNorm output (GEMM A operand)
//   qk       bf16 [B*T][2d],  vt bf16 [d][B*T]   fused QKV GEMM; V is written transposed for the P.V MMA
//   att      bf16 [B*T][d]              attention output
//   ff       bf16 [B*T][4d]             GELU(fc1)
// Every matrix product runs on the tcgen05 GEMM (gemm.cu) / attention kernel (attention.cu); LayerNorm and the
// mel re-layout are warp-shuffle / shared-memory-transpose kernels.
#include <algorithm>
#include "common.cuh"
#include "encoder.cuh"
#include "gemm.cuh"
#include "model.cuh"
#include "profile.cuh"

namespace wdr {

constexpr int kTPad = (WDR_AUDIO_CTX + 7) / 8 * 8;  // 1504: V^T column stride per window

int encoder_attention(const __nv_bfloat16* qk, const __nv_bfloat16* vt, int64_t ldt, int B, int T, int n_head, int d_model,
                      __nv_bfloat16* out, cudaStream_t st);
template <typename In>
int mel_launch(wdr_mel* m, const In* pcm, int64_t chunk_stride, const int32_t* n_valid_dev, int n_fixed, int n_chunks, int n_frames,
               int normalize, float* out, float* out_max, cudaStream_t st);

// raw log-mel [B][n_mel][n_frames] (+ per-window max) -> normalised bf16 frames [B][3002][128] (rows 1..3000)
// mel_offset selects the first frame (whisper_encode's mel offset); frames past n_frames read as the clamp floor 0?
// No: whisper.cpp zero-fills the encoder input beyond the mel length, so those rows are 0.
__global__ void mel_to_frames_kernel(const float* __restrict__ mel, int n_mel, int n_frames, int mel_offset,
                                     const float* __restrict__ chunk_max, int normalized_input, __nv_bfloat16* __restrict__ frames) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int f0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const float* m = mel + (int64_t)b * n_mel * n_frames;
    const float lo = normalized_input ? 0.0f : chunk_max[b] - 8.0f;
    const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
    for (int i = ty; i < 32; i += 8) {
        const int c = c0 + i, f = f0 + tx, src = mel_offset + f;
        float v = 0.0f;
        if (c < n_mel && f < WDR_CHUNK_FRAMES && src < n_frames) {
            v = m[(int64_t)c * n_frames + src];
            if (!normalized_input) v = (fmaxf(v, lo) + 4.0f) * 0.25f;
        }
        tile[i][tx] = v;
    }
    __syncthreads();
    __nv_bfloat16* o = frames + (int64_t)b * 3002 * kConv1CPad;
    for (int i = ty; i < 32; i += 8) {
        const int f = f0 + i, c = c0 + tx;
        if (f < WDR_CHUNK_FRAMES) o[(int64_t)(f + 1) * kConv1CPad + c] = __float2bfloat16_rn(tile[tx][i]);
    }
}

// LayerNorm over the last dim (eps 1e-5): one warp per row, fp32 in, bf16 or fp32 out.
template <typename Out>
__global__ void layernorm_kernel(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ bta, int64_t rows,
                                 int d, Out* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float4* xr = reinterpret_cast<const float4*>(x + row * d);
    const int nv = d >> 2;  // float4 per row; d % 128 == 0 -> nv % 32 == 0
    float4 v[10];           // d <= 1280
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < 10; i++) {
        if (i * 32 + lane < nv) {
            v[i] = xr[i * 32 + lane];
            s += v[i].x + v[i].y + v[i].z + v[i].w;
        }
    }
    s = warp_sum(s);
    const float mean = s / (float)d;
    float q = 0.0f;
#pragma unroll
    for (int i = 0; i < 10; i++) {
        if (i * 32 + lane < nv) {
            const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
            q += a * a + b * b + c * c + e * e;
        }
    }
    q = warp_sum(q);
    const float rstd = rsqrtf(q / (float)d + 1e-5f);
#pragma unroll
    for (int i = 0; i < 10; i++) {
        if (i * 32 + lane < nv) {
            const int c4 = i * 32 + lane;
            const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + c4);
            const float4 bb = __ldg(reinterpret_cast<const float4*>(bta) + c4);
            const float y0 = (v[i].x - mean) * rstd * gg.x + bb.x;
            const float y1 = (v[i].y - mean) * rstd * gg.y + bb.y;
            const float y2 = (v[i].z - mean) * rstd * gg.z + bb.z;
            const float y3 = (v[i].w - mean) * rstd * gg.w + bb.w;
            if (sizeof(Out) == 2) {
                __nv_bfloat162 p0 = __floats2bfloat162_rn(y0, y1), p1 = __floats2bfloat162_rn(y2, y3);
                uint2 w;
                w.x = *reinterpret_cast<uint32_t*>(&p0);
                w.y = *reinterpret_cast<uint32_t*>(&p1);
                reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out) + row * d)[c4] = w;
            } else {
                reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + row * d)[c4] = make_float4(y0, y1, y2, y3);
            }
        }
    }
}

template <typename Out>
int layernorm(const float* x, const float* g, const float* b, int64_t rows, int d, Out* out, cudaStream_t st) {
    WDR_REQUIRE(d % 128 == 0 && d <= 1280, "layernorm: d must be a multiple of 128 and <= 1280");
    const int warps = 8;
    layernorm_kernel<Out><<<(unsigned)((rows + warps - 1) / warps), warps * 32, 0, st>>>(x, g, b, rows, d, out);
    WDR_LAUNCH_CHECK();
    return WDR_OK;
}
template int layernorm<float>(const float*, const float*, const float*, int64_t, int, float*, cudaStream_t);
template int layernorm<__nv_bfloat16>(const float*, const float*, const float*, int64_t, int, __nv_bfloat16*, cudaStream_t);

void EncoderWorkspace::release() {
    for (void* p : {(void*)mel_raw, (void*)chunk_max, (void*)frames, (void*)conv1, (void*)x, (void*)h, (void*)qk, (void*)vt, (void*)att,
                    (void*)ff})
        if (p) cudaFree(p);
    *this = EncoderWorkspace();
}

int EncoderWorkspace::reserve(const WhisperArch& a, int B) {
    if (B <= cap) return WDR_OK;
    release();
    const int64_t M = (int64_t)B * WDR_AUDIO_CTX;
    const int d = a.d;
    ldt = (int64_t)B * kTPad;
    WDR_CUDA_TRY(cudaMalloc(&mel_raw, sizeof(float) * (size_t)B * a.n_mel * WDR_CHUNK_FRAMES));
    WDR_CUDA_TRY(cudaMalloc(&chunk_max, sizeof(float) * B));
    WDR_CUDA_TRY(cudaMalloc(&frames, sizeof(__nv_bfloat16) * (size_t)B * 3002 * kConv1CPad));
    WDR_CUDA_TRY(cudaMalloc(&conv1, sizeof(__nv_bfloat16) * (size_t)B * 3002 * d));
    WDR_CUDA_TRY(cudaMalloc(&x, sizeof(float) * (size_t)M * d));
    WDR_CUDA_TRY(cudaMalloc(&h, sizeof(__nv_bfloat16) * (size_t)M * d));
    WDR_CUDA_TRY(cudaMalloc(&qk, sizeof(__nv_bfloat16) * (size_t)M * 2 * d));
    WDR_CUDA_TRY(cudaMalloc(&vt, sizeof(__nv_bfloat16) * (size_t)d * ldt));
    WDR_CUDA_TRY(cudaMalloc(&att, sizeof(__nv_bfloat16) * (size_t)M * d));
    WDR_CUDA_TRY(cudaMalloc(&ff, sizeof(__nv_bfloat16) * (size_t)M * 4 * d));
    // pad rows (0 and 3001 of every window) must read as zero; everything else is overwritten per call
    WDR_CUDA_TRY(cudaMemset(frames, 0, sizeof(__nv_bfloat16) * (size_t)B * 3002 * kConv1CPad));
    WDR_CUDA_TRY(cudaMemset(conv1, 0, sizeof(__nv_bfloat16) * (size_t)B * 3002 * d));
    WDR_CUDA_TRY(cudaMemset(vt, 0, sizeof(__nv_bfloat16) * (size_t)d * ldt));
    cap = B;
    return WDR_OK;
}

// mel (raw or already normalised) -> hidden states.  mel: [B][n_mel][n_frames] device.
int encoder_forward(const wdr_context* ctx, EncoderWorkspace& ws, const float* mel, int n_frames, int mel_offset,
                    const float* chunk_max, int normalized_input, int B, float* out_f32, __nv_bfloat16* out_bf16, cudaStream_t st,
                    Profiler* prof) {
    const WhisperArch& a = ctx->arch;
    const WhisperWeights& w = ctx->w;
    const int d = a.d, T = WDR_AUDIO_CTX;
    const int64_t M = (int64_t)B * T;
    int rc;
    {
        ProfScope ps(prof, KC_MEL_AUX, st);
        dim3 grid((WDR_CHUNK_FRAMES + 31) / 32, kConv1CPad / 32, B);
        mel_to_frames_kernel<<<grid, dim3(32, 8), 0, st>>>(mel, a.n_mel, n_frames, mel_offset, chunk_max, normalized_input, ws.frames);
        WDR_LAUNCH_CHECK();
    }
    {   // conv1: k=3, stride 1, pad 1 -> rows i reads frames rows i, i+1, i+2 (row 0 is the left pad)
        GemmDesc g;
        g.A = ws.frames; g.a_row_stride = kConv1CPad; g.a_batch_stride = (int64_t)3002 * kConv1CPad;
        g.rows_per_batch = WDR_CHUNK_FRAMES; g.n_batch = B;
        g.W = w.conv1_w; g.ldw = 3 * kConv1CPad; g.N = d; g.K = 3 * kConv1CPad; g.kb_per_tap = kConv1CPad / 64; g.a_cols = kConv1CPad;
        g.epilogue = EPI_BIAS_GELU_BF16; g.out = ws.conv1 + d; g.ldc = d; g.c_batch_stride = (int64_t)3002 * d; g.bias = w.conv1_b;
        ProfScope ps(prof, KC_GEMM, st);
        if ((rc = gemm_bf16(g, st)) != WDR_OK) return rc;
    }
    {   // conv2: k=3, stride 2, pad 1 over row pairs: output j reads pair rows j (taps 0,1) and j+1 (tap 2)
        GemmDesc g;
        g.A = ws.conv1; g.a_row_stride = 2 * d; g.a_batch_stride = (int64_t)3002 * d;
        g.rows_per_batch = T; g.n_batch = B;
        g.W = w.conv2_w; g.ldw = 3 * d; g.N = d; g.K = 3 * d; g.kb_per_tap = 2 * d / 64; g.a_cols = 2 * d;
        g.epilogue = EPI_BIAS_GELU_POS_F32; g.out = ws.x; g.ldc = d; g.bias = w.conv2_b; g.pos = w.enc_pos;
        ProfScope ps(prof, KC_GEMM, st);
        if ((rc = gemm_bf16(g, st)) != WDR_OK) return rc;
    }
    for (int l = 0; l < a.n_enc_layer; l++) {
        const EncLayerW& e = w.enc[l];
        {
            ProfScope ps(prof, KC_LAYERNORM, st);
            if ((rc = layernorm<__nv_bfloat16>(ws.x, e.ln1_g, e.ln1_b, M, d, ws.h, st)) != WDR_OK) return rc;
        }
        {
            GemmDesc g;
            // per-window batches so that V^T columns can be laid out at a 16-byte aligned per-window stride
            g.A = ws.h; g.a_row_stride = d; g.a_batch_stride = (int64_t)T * d; g.rows_per_batch = T; g.n_batch = B;
            g.W = e.w_qkv; g.ldw = d; g.N = 3 * d; g.K = d;
            g.epilogue = EPI_QKV_BF16; g.out = ws.qk; g.ldc = 2 * d; g.bias = e.b_qkv;
            g.out_t = ws.vt; g.ldt = ws.ldt; g.t_batch_stride = kTPad; g.n_split = 2 * d;
            ProfScope ps(prof, KC_GEMM, st);
        if ((rc = gemm_bf16(g, st)) != WDR_OK) return rc;
        }
        {
            ProfScope ps(prof, KC_ATTENTION, st);
            if ((rc = encoder_attention(ws.qk, ws.vt, ws.ldt, B, T, a.n_head, d, ws.att, st)) != WDR_OK) return rc;
        }
        {
            GemmDesc g;
            g.A = ws.att; g.a_row_stride = d; g.rows_per_batch = (int)M; g.n_batch = 1;
            g.W = e.w_o; g.ldw = d; g.N = d; g.K = d;
            g.epilogue = EPI_BIAS_RESID_F32; g.out = ws.x; g.ldc = d; g.bias = e.b_o; g.resid = ws.x;
            ProfScope ps(prof, KC_GEMM, st);
        if ((rc = gemm_bf16(g, st)) != WDR_OK) return rc;
        }
        {
            ProfScope ps(prof, KC_LAYERNORM, st);
            if ((rc = layernorm<__nv_bfloat16>(ws.x, e.ln2_g, e.ln2_b, M, d, ws.h, st)) != WDR_OK) return rc;
        }
        {
            GemmDesc g;
            g.A = ws.h; g.a_row_stride = d; g.rows_per_batch = (int)M; g.n_batch = 1;
            g.W = e.w_fc1; g.ldw = d; g.N = 4 * d; g.K = d;
            g.epilogue = EPI_BIAS_GELU_BF16; g.out = ws.ff; g.ldc = 4 * d; g.bias = e.b_fc1;
            ProfScope ps(prof, KC_GEMM, st);
        if ((rc = gemm_bf16(g, st)) != WDR_OK) return rc;
        }
        {
            GemmDesc g;
            g.A = ws.ff; g.a_row_stride = 4 * d; g.rows_per_batch = (int)M; g.n_batch = 1;
            g.W = e.w_fc2; g.ldw = 4 * d; g.N = d; g.K = 4 * d;
            g.epilogue = EPI_BIAS_RESID_F32; g.out = ws.x; g.ldc = d; g.bias = e.b_fc2; g.resid = ws.x;
            ProfScope ps(prof, KC_GEMM, st);
        if ((rc = gemm_bf16(g, st)) != WDR_OK) return rc;
        }
    }
    ProfScope ps_post(prof, KC_LAYERNORM, st);
    if (out_f32 && (rc = layernorm<float>(ws.x, w.enc_lnpost_g, w.enc_lnpost_b, M, d, out_f32, st)) != WDR_OK) return rc;
    if (out_bf16 && (rc = layernorm<__nv_bfloat16>(ws.x, w.enc_lnpost_g, w.enc_lnpost_b, M, d, out_bf16, st)) != WDR_OK) return rc;
    return WDR_OK;
}

// PCM windows -> hidden states (mel + encoder), everything on `st`.
template <typename In>
int encode_chunks(const wdr_context* ctx, EncoderWorkspace& ws, const In* pcm, int64_t chunk_stride, const int32_t* n_valid_dev, int B,
                  float* out_f32, __nv_bfloat16* out_bf16, cudaStream_t st, Profiler* prof) {
    int rc = ws.reserve(ctx->arch, B);
    if (rc != WDR_OK) return rc;
    {
        ProfScope ps(prof, KC_MEL, st);
        rc = mel_launch<In>(ctx->mel, pcm, chunk_stride, n_valid_dev, WDR_CHUNK_SAMPLES, B, WDR_CHUNK_FRAMES, 0, ws.mel_raw, ws.chunk_max, st);
        if (rc != WDR_OK) return rc;
    }
    return encoder_forward(ctx, ws, ws.mel_raw, WDR_CHUNK_FRAMES, 0, ws.chunk_max, 0, B, out_f32, out_bf16, st, prof);
}
template int encode_chunks<int16_t>(const wdr_context*, EncoderWorkspace&, const int16_t*, int64_t, const int32_t*, int, float*, __nv_bfloat16*, cudaStream_t, Profiler*);
template int encode_chunks<float>(const wdr_context*, EncoderWorkspace&, const float*, int64_t, const int32_t*, int, float*, __nv_bfloat16*, cudaStream_t, Profiler*);

}  // namespace wdr
