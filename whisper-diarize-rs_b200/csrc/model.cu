// model.cu — Whisper architecture table, weight residency in HBM and the context (whisper_context) entry points.
//
// Replaces whisper_init_from_file_with_params / whisper_free (WhisperContext::new_with_params, reference
// src/transcribe.rs:102-155).  No model file can exist in this environment (no network), so the
// context synthesises weights of the named architecture from a seed with a counter-based generator that the CPU
// oracle reproduces bit for bit (uniform values, GEMM matrices rounded to bf16 once = "the checkpoint is bf16").
// Layout in HBM: every GEMM weight is [out][in] bf16 (K-major B operand for tcgen05), Q|K|V fused into one
// [3d][d] matrix, conv kernels stored tap-major ([out][tap][in]) for the implicit-GEMM stem.
#include <math.h>
#include <string.h>
#include <string>
#include "common.cuh"
#include "model.cuh"
#include "ggml_file.cuh"

namespace wdr {

static const WhisperArch kArchs[] = {
    {"tiny.en", 384, 6, 4, 4, 80, 51864, false, 0},   {"tiny", 384, 6, 4, 4, 80, 51865, true, 1},
    {"base.en", 512, 8, 6, 6, 80, 51864, false, 2},   {"base", 512, 8, 6, 6, 80, 51865, true, 3},
    {"small.en", 768, 12, 12, 12, 80, 51864, false, 4}, {"small", 768, 12, 12, 12, 80, 51865, true, 5},
    {"medium.en", 1024, 16, 24, 24, 80, 51864, false, 6}, {"medium", 1024, 16, 24, 24, 80, 51865, true, 7},
    {"large-v3", 1280, 20, 32, 32, 128, 51866, true, 8}, {"large-v3-turbo", 1280, 20, 32, 4, 128, 51866, true, 9},
};

// (layer, head) alignment heads per DTW preset == whisper_alignment_heads_preset (OpenAI _ALIGNMENT_HEADS)
static const std::vector<std::pair<int, int>> kAheads[10] = {
    {{1, 0}, {2, 0}, {2, 5}, {3, 0}, {3, 1}, {3, 2}, {3, 3}, {3, 4}},
    {{2, 2}, {3, 0}, {3, 2}, {3, 3}, {3, 4}, {3, 5}},
    {{3, 3}, {4, 7}, {5, 1}, {5, 5}, {5, 7}},
    {{3, 1}, {4, 2}, {4, 3}, {4, 7}, {5, 1}, {5, 2}, {5, 4}, {5, 6}},
    {{6, 6}, {7, 0}, {7, 3}, {7, 8}, {8, 2}, {8, 5}, {8, 7}, {9, 0}, {9, 4}, {9, 8}, {9, 10}, {10, 0}, {10, 1}, {10, 2}, {10, 3}, {10, 6}, {10, 11}, {11, 2}, {11, 4}},
    {{5, 3}, {5, 9}, {8, 0}, {8, 4}, {8, 7}, {8, 8}, {9, 0}, {9, 7}, {9, 9}, {10, 5}},
    {{11, 4}, {14, 1}, {14, 12}, {14, 14}, {15, 4}, {16, 0}, {16, 4}, {16, 9}, {17, 12}, {17, 14}, {18, 7}, {18, 10}, {18, 15}, {20, 0}, {20, 3}, {20, 9}, {20, 14}, {21, 12}},
    {{13, 15}, {15, 4}, {15, 15}, {16, 1}, {20, 0}, {23, 4}},
    {{7, 0}, {10, 17}, {12, 18}, {13, 12}, {16, 1}, {17, 14}, {19, 11}, {21, 4}, {24, 1}, {25, 6}},
    {{2, 4}, {2, 11}, {3, 3}, {3, 6}, {3, 11}, {3, 14}},
};
const std::vector<std::pair<int, int>>* aheads_for_preset(int preset) {
    return (preset >= 0 && preset < 10) ? &kAheads[preset] : nullptr;
}

const WhisperArch* find_arch(const char* name) {
    if (!name) return nullptr;
    for (const auto& a : kArchs)
        if (strcmp(a.name, name) == 0) return &a;
    return nullptr;
}

// ---------------------------------------------------------------------------------------------------
// counter-based weight synthesis (the CPU checker in the test tree restates it bit for bit)
// ---------------------------------------------------------------------------------------------------
__host__ __device__ inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__host__ __device__ inline float synth_unit(uint64_t key, uint64_t i) {
    const uint64_t z = splitmix64(key + i);
    const int k = (int)(z >> 40);  // 24 bits
    return (float)(k - 8388608) * (1.0f / 8388608.0f);  // [-1, 1), exact in fp32
}
static uint64_t fnv1a(const char* s) {
    uint64_t h = 0xcbf29ce484222325ULL;
    for (; *s; s++) { h ^= (unsigned char)*s; h *= 0x100000001b3ULL; }
    return h;
}
static uint64_t tensor_key(uint64_t seed, const std::string& name) { return splitmix64(fnv1a(name.c_str()) ^ splitmix64(seed)); }

// mode 0: dst[r][c] <- canonical r*cols + c.  mode 1 (conv, canonical [out][cin][3]): dst[o][t*cpad + c] <- (o*cin + c)*3 + t
template <typename T>
__global__ void synth_kernel(T* __restrict__ dst, int rows, int cols, int64_t ld, uint64_t key, float offset, float scale, int mode,
                             int cin, int cpad) {
    const int64_t total = (int64_t)rows * cols;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(e / cols), c = (int)(e % cols);
        float v;
        if (mode == 0) {
            v = __fadd_rn(offset, __fmul_rn(synth_unit(key, (uint64_t)e), scale));  // no FMA contraction: matches the checker
        } else {
            const int t = c / cpad, ci = c % cpad;
            v = (ci < cin) ? __fadd_rn(offset, __fmul_rn(synth_unit(key, ((uint64_t)r * cin + ci) * 3 + t), scale)) : 0.0f;
        }
        if (sizeof(T) == 2) {
            // matrices are bf16 "checkpoints": round-to-nearest-even once
            reinterpret_cast<__nv_bfloat16*>(dst)[(int64_t)r * ld + c] = __float2bfloat16_rn(v);
        } else {
            reinterpret_cast<float*>(dst)[(int64_t)r * ld + c] = v;
        }
    }
}

// [N][K] row-major -> k-block-major [K/64][N][64]: for a fixed 64-wide k-block the rows of a weight tile are contiguous, so the
// TMA box of a weight-streaming GEMM (BN rows x 128 B) is ONE contiguous BN*128-byte read instead of BN pieces at a K*2-byte stride.
__global__ void kb_major_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst, int N, int K) {
    const int64_t total = (int64_t)N * K / 8;  // 16-byte vectors
    const int vec_per_kb = 8, num_kb = K / 64;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (int64_t)gridDim.x * blockDim.x) {
        const int64_t o = v / vec_per_kb;            // (kb, n) pair index in the destination
        const int c8 = (int)(v - o * vec_per_kb);
        const int kb = (int)(o / N), n = (int)(o - (int64_t)kb * N);
        reinterpret_cast<uint4*>(dst)[v] = reinterpret_cast<const uint4*>(src)[((int64_t)n * K + kb * 64) / 8 + c8];
    }
    (void)num_kb;
}

// file-backed twin of synth_kernel: same destination mapping, values read from the tensor's canonical (OpenAI) layout
template <typename T>
__global__ void load_kernel(T* __restrict__ dst, int rows, int cols, int64_t ld, const float* __restrict__ src, int mode, int cin, int cpad) {
    const int64_t total = (int64_t)rows * cols;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(e / cols), c = (int)(e % cols);
        float v;
        if (mode == 0) v = src[e];
        else {
            const int t = c / cpad, ci = c % cpad;
            v = (ci < cin) ? src[((int64_t)r * cin + ci) * 3 + t] : 0.0f;
        }
        if (sizeof(T) == 2) reinterpret_cast<__nv_bfloat16*>(dst)[(int64_t)r * ld + c] = __float2bfloat16_rn(v);
        else reinterpret_cast<float*>(dst)[(int64_t)r * ld + c] = v;
    }
}

struct Builder {
    wdr_context* ctx;
    int rc = WDR_OK;
    const GgmlFile* file = nullptr;  // non-null: tensors come from a ggml checkpoint instead of the seeded generator
    template <typename T>
    T* alloc(size_t n) {
        if (rc != WDR_OK) return nullptr;
        void* p = nullptr;
        cudaError_t e = cudaMalloc(&p, n * sizeof(T));
        if (e != cudaSuccess) {
            set_error("weights: cudaMalloc(%zu) -> %s", n * sizeof(T), cudaGetErrorString(e));
            rc = WDR_ERR_OOM;
            return nullptr;
        }
        ctx->allocations.push_back(p);
        ctx->weight_bytes += n * sizeof(T);
        return reinterpret_cast<T*>(p);
    }
    template <typename T>
    void fill(T* dst, int rows, int cols, int64_t ld, const std::string& name, float offset, float scale, int mode = 0, int cin = 0, int cpad = 1) {
        if (rc != WDR_OK) return;
        const int64_t total = (int64_t)rows * cols;
        int blocks = (int)((total + 255) / 256);
        if (blocks > 148 * 8) blocks = 148 * 8;
        if (file) {
            std::vector<float> host;
            std::string err;
            const int64_t expect = mode == 0 ? total : (int64_t)rows * cin * 3;
            if (!file->read_f32(name, expect, &host, &err)) { set_error("%s", err.c_str()); rc = WDR_ERR_INVALID; return; }
            float* tmp = nullptr;
            if (cudaMalloc(&tmp, sizeof(float) * host.size()) != cudaSuccess) { set_error("weights: staging buffer for '%s'", name.c_str()); rc = WDR_ERR_OOM; return; }
            cudaMemcpy(tmp, host.data(), sizeof(float) * host.size(), cudaMemcpyHostToDevice);
            load_kernel<T><<<blocks, 256>>>(dst, rows, cols, ld, tmp, mode, cin, cpad);
            count_launch();
            cudaDeviceSynchronize();
            cudaFree(tmp);
            if (cudaGetLastError() != cudaSuccess) { set_error("weight load kernel failed for '%s'", name.c_str()); rc = WDR_ERR_CUDA; }
            return;
        }
        synth_kernel<T><<<blocks, 256>>>(dst, rows, cols, ld, tensor_key(ctx->seed, name), offset, scale, mode, cin, cpad);
        count_launch();
        if (cudaGetLastError() != cudaSuccess) { set_error("synth kernel launch failed"); rc = WDR_ERR_CUDA; }
    }
    // in-place re-layout of a finished [N][K] matrix to k-block-major (decoder weights: every consumer is a weight-streaming GEMM)
    void to_kb_major(__nv_bfloat16* w, int N, int K) {
        if (rc != WDR_OK) return;
        if (K % 64 != 0 || N % 8 != 0) { set_error("kb-major layout needs K %% 64 == 0"); rc = WDR_ERR_INVALID; return; }
        __nv_bfloat16* tmp = nullptr;
        if (cudaMalloc(&tmp, sizeof(__nv_bfloat16) * (size_t)N * K) != cudaSuccess) { set_error("weights: temporary for the kb-major layout"); rc = WDR_ERR_OOM; return; }
        kb_major_kernel<<<148 * 8, 256>>>(w, tmp, N, K);
        count_launch();
        cudaMemcpy(w, tmp, sizeof(__nv_bfloat16) * (size_t)N * K, cudaMemcpyDeviceToDevice);
        cudaFree(tmp);
        if (cudaGetLastError() != cudaSuccess) { set_error("kb-major re-layout failed"); rc = WDR_ERR_CUDA; }
    }
    __nv_bfloat16* matrix(int rows, int cols, const std::string& name, float scale) {
        auto* p = alloc<__nv_bfloat16>((size_t)rows * cols);
        fill(p, rows, cols, cols, name, 0.0f, scale);
        return p;
    }
    float* vec(int n, const std::string& name, float offset, float scale) {
        auto* p = alloc<float>(n);
        fill(p, 1, n, n, name, offset, scale);
        return p;
    }
    float* zeros(int n) {
        auto* p = alloc<float>(n);
        if (p) cudaMemset(p, 0, sizeof(float) * n);
        return p;
    }
};

static const float kWScale = 0.034641016151377546f;  // uniform half-width with std 0.02
static const float kBScale = 0.02f;

static int build_weights(wdr_context* ctx, const GgmlFile* file = nullptr) {
    const WhisperArch& a = ctx->arch;
    WhisperWeights& w = ctx->w;
    Builder b{ctx};
    b.file = file;
    const int d = a.d;
    // ---- encoder ----
    w.conv1_w = b.alloc<__nv_bfloat16>((size_t)d * 3 * kConv1CPad);
    b.fill(w.conv1_w, d, 3 * kConv1CPad, 3 * kConv1CPad, "encoder.conv1.weight", 0.0f, kWScale, 1, a.n_mel, kConv1CPad);
    w.conv1_b = b.vec(d, "encoder.conv1.bias", 0.0f, kBScale);
    w.conv2_w = b.alloc<__nv_bfloat16>((size_t)d * 3 * d);
    b.fill(w.conv2_w, d, 3 * d, 3 * d, "encoder.conv2.weight", 0.0f, kWScale, 1, d, d);
    w.conv2_b = b.vec(d, "encoder.conv2.bias", 0.0f, kBScale);
    if (file) {
        w.enc_pos = b.vec(WDR_AUDIO_CTX * d, "encoder.positional_embedding", 0.0f, 0.0f);
    } else {
        // sinusoids(1500, d) of OpenAI Whisper, evaluated in fp32
        std::vector<float> pos((size_t)WDR_AUDIO_CTX * d);
        const int half = d / 2;
        const float inc = logf(10000.0f) / (float)(half - 1);
        for (int t = 0; t < WDR_AUDIO_CTX; t++)
            for (int i = 0; i < half; i++) {
                const float st = (float)t * expf(-inc * (float)i);
                pos[(size_t)t * d + i] = sinf(st);
                pos[(size_t)t * d + half + i] = cosf(st);
            }
        w.enc_pos = b.alloc<float>(pos.size());
        if (w.enc_pos) cudaMemcpy(w.enc_pos, pos.data(), pos.size() * sizeof(float), cudaMemcpyHostToDevice);
    }
    const float enc_out_scale = kWScale / sqrtf(2.0f * a.n_enc_layer);
    w.enc.resize(a.n_enc_layer);
    for (int l = 0; l < a.n_enc_layer; l++) {
        const std::string p = "encoder.blocks." + std::to_string(l) + ".";
        EncLayerW& e = w.enc[l];
        e.ln1_g = b.vec(d, p + "attn_ln.weight", 1.0f, 0.1f);
        e.ln1_b = b.vec(d, p + "attn_ln.bias", 0.0f, 0.1f);
        e.w_qkv = b.alloc<__nv_bfloat16>((size_t)3 * d * d);
        b.fill(e.w_qkv, d, d, d, p + "attn.query.weight", 0.0f, kWScale);
        b.fill(e.w_qkv + (size_t)d * d, d, d, d, p + "attn.key.weight", 0.0f, kWScale);
        b.fill(e.w_qkv + (size_t)2 * d * d, d, d, d, p + "attn.value.weight", 0.0f, kWScale);
        e.b_qkv = b.zeros(3 * d);
        b.fill(e.b_qkv, 1, d, d, p + "attn.query.bias", 0.0f, kBScale);
        b.fill(e.b_qkv + 2 * d, 1, d, d, p + "attn.value.bias", 0.0f, kBScale);
        e.w_o = b.matrix(d, d, p + "attn.out.weight", enc_out_scale);
        e.b_o = b.vec(d, p + "attn.out.bias", 0.0f, kBScale);
        e.ln2_g = b.vec(d, p + "mlp_ln.weight", 1.0f, 0.1f);
        e.ln2_b = b.vec(d, p + "mlp_ln.bias", 0.0f, 0.1f);
        e.w_fc1 = b.matrix(4 * d, d, p + "mlp.0.weight", kWScale);
        e.b_fc1 = b.vec(4 * d, p + "mlp.0.bias", 0.0f, kBScale);
        e.w_fc2 = b.matrix(d, 4 * d, p + "mlp.2.weight", enc_out_scale);
        e.b_fc2 = b.vec(d, p + "mlp.2.bias", 0.0f, kBScale);
    }
    w.enc_lnpost_g = b.vec(d, "encoder.ln_post.weight", 1.0f, 0.1f);
    w.enc_lnpost_b = b.vec(d, "encoder.ln_post.bias", 0.0f, 0.1f);
    // ---- decoder ----
    {
        const int n_pad = (a.n_vocab + 7) / 8 * 8;
        w.tok_emb = b.alloc<__nv_bfloat16>((size_t)n_pad * d);
        if (w.tok_emb) cudaMemset(w.tok_emb, 0, sizeof(__nv_bfloat16) * (size_t)n_pad * d);
        b.fill(w.tok_emb, a.n_vocab, d, d, "decoder.token_embedding.weight", 0.0f, kWScale);
    }
    w.dec_pos = b.vec(WDR_TEXT_CTX * d, "decoder.positional_embedding", 0.0f, 0.017320508f);
    const float dec_out_scale = kWScale / sqrtf(2.0f * a.n_dec_layer);
    w.dec.resize(a.n_dec_layer);
    for (int l = 0; l < a.n_dec_layer; l++) {
        const std::string p = "decoder.blocks." + std::to_string(l) + ".";
        DecLayerW& e = w.dec[l];
        e.ln1_g = b.vec(d, p + "attn_ln.weight", 1.0f, 0.1f);
        e.ln1_b = b.vec(d, p + "attn_ln.bias", 0.0f, 0.1f);
        e.w_qkv = b.alloc<__nv_bfloat16>((size_t)3 * d * d);
        b.fill(e.w_qkv, d, d, d, p + "attn.query.weight", 0.0f, kWScale);
        b.fill(e.w_qkv + (size_t)d * d, d, d, d, p + "attn.key.weight", 0.0f, kWScale);
        b.fill(e.w_qkv + (size_t)2 * d * d, d, d, d, p + "attn.value.weight", 0.0f, kWScale);
        e.b_qkv = b.zeros(3 * d);
        b.fill(e.b_qkv, 1, d, d, p + "attn.query.bias", 0.0f, kBScale);
        b.fill(e.b_qkv + 2 * d, 1, d, d, p + "attn.value.bias", 0.0f, kBScale);
        e.w_o = b.matrix(d, d, p + "attn.out.weight", dec_out_scale);
        e.b_o = b.vec(d, p + "attn.out.bias", 0.0f, kBScale);
        e.ln2_g = b.vec(d, p + "cross_attn_ln.weight", 1.0f, 0.1f);
        e.ln2_b = b.vec(d, p + "cross_attn_ln.bias", 0.0f, 0.1f);
        e.w_cq = b.matrix(d, d, p + "cross_attn.query.weight", kWScale);
        e.b_cq = b.vec(d, p + "cross_attn.query.bias", 0.0f, kBScale);
        e.w_ckv = b.alloc<__nv_bfloat16>((size_t)2 * d * d);
        b.fill(e.w_ckv, d, d, d, p + "cross_attn.key.weight", 0.0f, kWScale);
        b.fill(e.w_ckv + (size_t)d * d, d, d, d, p + "cross_attn.value.weight", 0.0f, kWScale);
        e.b_ckv = b.zeros(2 * d);
        b.fill(e.b_ckv + d, 1, d, d, p + "cross_attn.value.bias", 0.0f, kBScale);
        e.w_co = b.matrix(d, d, p + "cross_attn.out.weight", dec_out_scale);
        e.b_co = b.vec(d, p + "cross_attn.out.bias", 0.0f, kBScale);
        e.ln3_g = b.vec(d, p + "mlp_ln.weight", 1.0f, 0.1f);
        e.ln3_b = b.vec(d, p + "mlp_ln.bias", 0.0f, 0.1f);
        e.w_fc1 = b.matrix(4 * d, d, p + "mlp.0.weight", kWScale);
        e.b_fc1 = b.vec(4 * d, p + "mlp.0.bias", 0.0f, kBScale);
        e.w_fc2 = b.matrix(d, 4 * d, p + "mlp.2.weight", dec_out_scale);
        e.b_fc2 = b.vec(d, p + "mlp.2.bias", 0.0f, kBScale);
        // the decode-time consumers of these six matrices are weight-streaming GEMMs (GemmDesc::w_kb_major)
        b.to_kb_major(e.w_qkv, 3 * d, d);
        b.to_kb_major(e.w_o, d, d);
        b.to_kb_major(e.w_cq, d, d);
        b.to_kb_major(e.w_co, d, d);
        b.to_kb_major(e.w_fc1, 4 * d, d);
        b.to_kb_major(e.w_fc2, d, 4 * d);
    }
    w.dec_ln_g = b.vec(d, "decoder.ln.weight", 1.0f, 0.1f);
    w.dec_ln_b = b.vec(d, "decoder.ln.bias", 0.0f, 0.1f);
    if (b.rc != WDR_OK) return b.rc;
    WDR_CUDA_TRY(cudaDeviceSynchronize());
    return WDR_OK;
}

// librosa.filters.mel(sr=16000, n_fft=400, n_mels, fmin=0, fmax=8000, htk=False, norm="slaney") — the matrix
// whisper.cpp reads from the ggml model file.
static double hz_to_mel_slaney(double f) {
    const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = log(6.4) / 27.0;
    return f >= min_log_hz ? min_log_mel + log(f / min_log_hz) / logstep : f / f_sp;
}
static double mel_to_hz_slaney(double m) {
    const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = log(6.4) / 27.0;
    return m >= min_log_mel ? min_log_hz * exp(logstep * (m - min_log_mel)) : f_sp * m;
}
void whisper_mel_filters(int n_mel, float* out /* [n_mel][201] */) {
    const int n_bins = 201;
    const double fmax = 8000.0;
    std::vector<double> hz(n_mel + 2);
    const double m_lo = hz_to_mel_slaney(0.0), m_hi = hz_to_mel_slaney(fmax);
    for (int i = 0; i < n_mel + 2; i++) hz[i] = mel_to_hz_slaney(m_lo + (m_hi - m_lo) * i / (n_mel + 1));
    for (int i = 0; i < n_mel; i++) {
        const double enorm = 2.0 / (hz[i + 2] - hz[i]);
        for (int k = 0; k < n_bins; k++) {
            const double f = fmax * k / (n_bins - 1);
            const double lower = (f - hz[i]) / (hz[i + 1] - hz[i]);
            const double upper = (hz[i + 2] - f) / (hz[i + 2] - hz[i + 1]);
            double v = lower < upper ? lower : upper;
            if (v < 0) v = 0;
            out[(size_t)i * n_bins + k] = (float)(v * enorm);
        }
    }
}

}  // namespace wdr

using namespace wdr;

extern "C" int wdr_mel_filters(int n_mel, float* out) {
    clear_error();
    WDR_REQUIRE(out && n_mel > 0 && n_mel <= 512, "bad arguments");
    whisper_mel_filters(n_mel, out);
    return WDR_OK;
}

extern "C" wdr_context_params wdr_context_default_params(void) {
    wdr_context_params p;
    memset(&p, 0, sizeof(p));
    p.use_gpu = 1;
    p.gpu_device = 0;
    p.flash_attn = 0;
    p.dtw_token_timestamps = 0;
    p.dtw_aheads_preset = WDR_AHEADS_NONE;
    p.dtw_mem_size = 128u * 1024 * 1024;
    p.arch_name = nullptr;
    p.seed = 1234;
    return p;
}

// WhisperArch of a checkpoint: the named architecture with the same dimensions when there is one (it carries the DTW preset), else
// an anonymous one.  Refuses geometries the kernels do not cover (head width != 64, contexts other than 1500 / 448).
static bool arch_from_file(const GgmlFile& f, WhisperArch* out, std::string* name, std::string* err) {
    if (f.n_audio_ctx != WDR_AUDIO_CTX || f.n_text_ctx != WDR_TEXT_CTX) { *err = "unsupported context sizes (need n_audio_ctx 1500, n_text_ctx 448)"; return false; }
    if (f.n_text_state != f.n_audio_state || f.n_text_head != f.n_audio_head || f.n_audio_state != 64 * f.n_audio_head) {
        *err = "unsupported geometry (need n_text_state == n_audio_state == 64 * n_head)";
        return false;
    }
    if (f.n_vocab < 51864 || f.n_vocab > 51866 || (f.n_mels != 80 && f.n_mels != 128)) { *err = "unsupported vocabulary / mel size"; return false; }
    if (f.filt_n_mel != f.n_mels || f.filt_n_fft != 201) { *err = "mel filterbank block does not match n_mels x 201"; return false; }
    for (const auto& a : kArchs)
        if (a.d == f.n_audio_state && a.n_head == f.n_audio_head && a.n_enc_layer == f.n_audio_layer && a.n_dec_layer == f.n_text_layer &&
            a.n_mel == f.n_mels && a.n_vocab == f.n_vocab) {
            *out = a;
            *name = a.name;
            return true;
        }
    WhisperArch a{};
    a.d = f.n_audio_state; a.n_head = f.n_audio_head; a.n_enc_layer = f.n_audio_layer; a.n_dec_layer = f.n_text_layer;
    a.n_mel = f.n_mels; a.n_vocab = f.n_vocab; a.multilingual = f.n_vocab >= 51865; a.dtw_preset = -1;
    *out = a;
    *name = "ggml-file";
    return true;
}

extern "C" wdr_context* wdr_init_from_file_with_params(const char* path, wdr_context_params params) {
    clear_error();
    if (!params.use_gpu) { set_error("use_gpu = false: libwdr_b200 has no CPU path"); return nullptr; }
    GgmlFile file;
    const bool from_file = path && path[0];
    WhisperArch arch_buf{};
    std::string arch_name;
    const WhisperArch* a = nullptr;
    if (from_file) {  // ggml-<model>.bin, as WhisperContext::new_with_params(model_path, ..) takes it (src/transcribe.rs:154)
        std::string err;
        if (!file.open(path, &err)) { set_error("%s", err.c_str()); return nullptr; }
        if (!arch_from_file(file, &arch_buf, &arch_name, &err)) { set_error("%s: %s", path, err.c_str()); return nullptr; }
        a = &arch_buf;
    } else {
        a = find_arch(params.arch_name);
        if (!a) { set_error("unknown architecture '%s'", params.arch_name ? params.arch_name : "(null)"); return nullptr; }
        arch_name = a->name;
    }
    if (ensure_device(params.gpu_device) != WDR_OK) return nullptr;
    wdr_context* ctx = new wdr_context();
    ctx->device = params.gpu_device;
    ctx->arch = *a;
    ctx->arch_name = arch_name;
    ctx->arch.name = ctx->arch_name.c_str();
    ctx->seed = params.seed;
    ctx->dtw_enabled = params.dtw_token_timestamps;
    ctx->dtw_preset = params.dtw_aheads_preset;
    ctx->dtw_mem_size = params.dtw_mem_size;
    ctx->flash_attn = params.flash_attn;
    if (params.dtw_token_timestamps) {
        // WDR_AHEADS_NONE with DTW on: fall back to the architecture's own preset (the crate always passes a preset, src/transcribe.rs:117-129)
        const int preset = params.dtw_aheads_preset >= 0 ? params.dtw_aheads_preset : a->dtw_preset;
        const auto* ah = aheads_for_preset(preset);
        if (!ah) { set_error("unknown alignment-head preset %d", preset); delete ctx; return nullptr; }
        for (auto& lh : *ah) {
            if (lh.first >= a->n_dec_layer || lh.second >= a->n_head) { set_error("alignment-head preset %d does not fit %s", preset, ctx->arch.name); delete ctx; return nullptr; }
            ctx->aheads.push_back(lh);
        }
        ctx->dtw_preset = preset;
    }
    if (build_weights(ctx, from_file ? &file : nullptr) != WDR_OK) { wdr_free(ctx); return nullptr; }
    if (from_file) {
        ctx->mel_filters = file.filters;      // whisper.cpp takes the filterbank from the checkpoint
        ctx->file_tokens = file.tokens;       // and the token strings
    } else {
        ctx->mel_filters.resize((size_t)a->n_mel * 201);
        whisper_mel_filters(a->n_mel, ctx->mel_filters.data());
    }
    ctx->mel = wdr_mel_init(ctx->mel_filters.data(), a->n_mel, ctx->device);
    if (!ctx->mel) { wdr_free(ctx); return nullptr; }
    log_msg(1, "wdr: %s d=%d heads=%d enc=%d dec=%d mel=%d vocab=%d weights=%.1f MB (seed %llu)", ctx->arch.name, a->d, a->n_head,
            a->n_enc_layer, a->n_dec_layer, a->n_mel, a->n_vocab, ctx->weight_bytes / 1048576.0, (unsigned long long)ctx->seed);
    return ctx;
}

extern "C" int wdr_ggml_probe(const char* path, int32_t* hparams, int32_t* n_tensors, int32_t* n_tokens) {
    clear_error();
    WDR_REQUIRE(path && hparams, "bad arguments");
    GgmlFile f;
    std::string err;
    if (!f.open(path, &err)) { set_error("%s", err.c_str()); return WDR_ERR_INVALID; }
    const int32_t hp[11] = {f.n_vocab, f.n_audio_ctx, f.n_audio_state, f.n_audio_head, f.n_audio_layer, f.n_text_ctx, f.n_text_state, f.n_text_head,
                            f.n_text_layer, f.n_mels, f.ftype};
    memcpy(hparams, hp, sizeof(hp));
    if (n_tensors) *n_tensors = (int32_t)f.tensors.size();
    if (n_tokens) *n_tokens = (int32_t)f.tokens.size();
    return WDR_OK;
}

extern "C" void wdr_free(wdr_context* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->mel) wdr_mel_free(ctx->mel);
    for (void* p : ctx->allocations) cudaFree(p);
    delete ctx;
    wdr::devbuf_trim();  // temporaries cached behind the host-pointer entry points go back to the driver with the model
}

extern "C" int wdr_model_info(const wdr_context* ctx, wdr_model_dims* out) {
    clear_error();
    WDR_REQUIRE(ctx && out, "null pointer");
    out->n_audio_state = ctx->arch.d;
    out->n_audio_head = ctx->arch.n_head;
    out->n_audio_layer = ctx->arch.n_enc_layer;
    out->n_text_layer = ctx->arch.n_dec_layer;
    out->n_mels = ctx->arch.n_mel;
    out->n_vocab = ctx->arch.n_vocab;
    out->n_audio_ctx = WDR_AUDIO_CTX;
    out->n_text_ctx = WDR_TEXT_CTX;
    out->is_multilingual = ctx->arch.multilingual ? 1 : 0;
    out->weight_bytes = (int64_t)ctx->weight_bytes;
    return WDR_OK;
}
