// full.cu — whisper_full_with_state as the reference drives it, batched over independent <= 30 s buffers.
//
// Replaces `state.full(params, &samples)` (reference src/transcribe.rs:389, parameters from setup_params :20-87) and the
// result accessors (`full_n_segments`, segment text/t0/t1, `n_tokens`, `get_token`, `token_data`; :393-412, :252-282).
// Device work: log-mel -> encoder -> cross-KV -> greedy decode (decoder.cu) -> DTW pass (teacher-forced decoder steps that
// capture the alignment heads, dtw.cu).  Host work (scalar, data dependent, restated from SURVEY A.4-A.6): segment assembly,
// whisper_exp_compute_token_level_timestamps, stamping t_dtw from the DTW path.
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <mutex>
#include <string>
#include <thread>
#include <map>
#include <random>
#include <vector>
#include <cuda_profiler_api.h>
#include "common.cuh"
#include "decoder.cuh"
#include "encoder.cuh"
#include "state.cuh"
#include "vocab.cuh"

namespace wdr {

template <typename In>
int mel_launch(wdr_mel* m, const In* pcm, int64_t chunk_stride, const int32_t* n_valid_dev, int n_fixed, int n_chunks, int n_frames, int normalize,
               float* out, float* out_max, cudaStream_t st);
int dtw_cost_dev(const float* w, int H, int T, int A, int sot_len, int width, float* mean, float* scale, float* out, cudaStream_t st);
struct DtwWindow { int64_t x_off; int32_t N, M; int64_t tr_off; };
int dtw_run(const float* x, std::vector<DtwWindow>& wins, int32_t* text_idx, int32_t* time_idx, int32_t* path_len, int max_path,
            float* cost_out, int32_t* trace_out, cudaStream_t st, void* wins_scratch_dev);

constexpr int kDeltaMin = 10;  // whisper.cpp v1.7.x: "if only 100 ms left, then stop" (delta_min = 10 mel frames)

// get_signal_energy: moving average of |x| over [i-hw, i+hw] clipped to the buffer, summed in index order
template <typename In>
__global__ void energy_kernel(const In* __restrict__ pcm, int64_t chunk_stride, const int32_t* __restrict__ n_valid, int n_fixed, int hw,
                              float* __restrict__ out, int64_t out_stride) {
    const int b = blockIdx.y;
    const int n = n_valid ? n_valid[b] : n_fixed;
    const In* x = pcm + (int64_t)b * chunk_stride;
    float* o = out + (int64_t)b * out_stride;
    const float den = (float)(2 * hw + 1);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float s = 0.0f;
        const int lo = max(0, i - hw), hi = min(n - 1, i + hw);
        for (int j = lo; j <= hi; j++) {
            float v;
            if (sizeof(In) == 2) v = (float)x[j] * (1.0f / 32768.0f); else v = (float)x[j];
            s += fabsf(v);
        }
        o[i] = s / den;
    }
}

__global__ void f32_to_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = __float2bfloat16_rn(in[i]);
}

static int timestamp_to_sample(int64_t t, int n_samples) {
    int64_t s = (t * WDR_SAMPLE_RATE) / 100;
    if (s > n_samples - 1) s = n_samples - 1;
    if (s < 0) s = 0;
    return (int)s;
}
static int64_t sample_to_timestamp(int i) { return (100ll * i) / WDR_SAMPLE_RATE; }

// f(i) for i in [0, n) on up to 16 host threads (dynamic claim); f must not touch shared state.  WDR_HOST_THREADS caps the count:
// with one process per GPU on a box with few cores, eight ranks x 16 threads would oversubscribe the host under the DTW pass.
template <typename F>
static void parallel_for(int n, F f) {
    static const int cap = getenv("WDR_HOST_THREADS") ? std::max(1, atoi(getenv("WDR_HOST_THREADS"))) : 16;
    int nt = (int)std::thread::hardware_concurrency();
    nt = std::max(1, std::min(std::min(nt, cap), n));
    if (nt <= 1) { for (int i = 0; i < n; i++) f(i); return; }
    std::atomic<int> next{0};
    std::vector<std::thread> th;
    auto work = [&]() { for (int i; (i = next.fetch_add(1)) < n;) f(i); };
    for (int t = 1; t < nt; t++) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
}

// whisper_exp_compute_token_level_timestamps for one segment; st3 = {t_beg, t_last, tid_last}.
// Part 1: everything up to the energy pass (timestamp-token anchors, voice-length split, monotonic fix-up) — scalar logic on the
// token list alone.  Returns true when the energy pass applies (n >= 2).
static bool token_level_timestamps_part1(std::vector<wdr_token_data>& tokens, int64_t t0, int64_t t1, const Vocab& v,
                                         int n_samples, float thold_pt, float thold_ptsum, int64_t* st3) {
    const int n = (int)tokens.size();
    if (n_samples == 0 || n == 0) return false;
    if (n == 1) { tokens[0].t0 = t0; tokens[0].t1 = t1; return false; }
    int64_t &t_beg = st3[0], &t_last = st3[1], &tid_last = st3[2];
    for (int j = 0; j < n; j++) {
        wdr_token_data& tk = tokens[j];
        if (j == 0) {
            if (tk.id == v.beg) {
                tokens[j].t0 = t0; tokens[j].t1 = t0; tokens[j + 1].t0 = t0;
                t_beg = t0; t_last = t0; tid_last = v.beg;
            } else {
                tokens[j].t0 = t_last;
            }
        }
        const int64_t tt = t_beg + 2 * (tk.tid - v.beg);
        tk.vlen = voice_length(token_text(v, tk.id));
        if (tk.pt > thold_pt && tk.ptsum > thold_ptsum && tk.tid > tid_last && tt <= t1) {
            if (j > 0) tokens[j - 1].t1 = tt;
            tokens[j].t0 = tt;
            tid_last = tk.tid;
        }
    }
    tokens[n - 2].t1 = t1; tokens[n - 1].t0 = t1; tokens[n - 1].t1 = t1;
    t_last = t1;
    {
        int p0 = 0, p1 = 0;
        while (true) {
            while (p1 < n && tokens[p1].t1 < 0) p1++;
            if (p1 >= n) p1--;
            if (p1 > p0) {
                double psum = 0.0;
                for (int j = p0; j <= p1; j++) psum += tokens[j].vlen;
                const double dt = (double)(tokens[p1].t1 - tokens[p0].t0);
                for (int j = p0 + 1; j <= p1; j++) {
                    const double ct = tokens[j - 1].t0 + dt * tokens[j - 1].vlen / psum;
                    tokens[j - 1].t1 = (int64_t)ct;
                    tokens[j].t0 = (int64_t)ct;
                }
            }
            p1++;
            p0 = p1;
            if (p1 >= n) break;
        }
    }
    for (int j = 0; j < n - 1; j++) {
        if (tokens[j].t1 < 0) tokens[j + 1].t0 = tokens[j].t1;
        if (j > 0 && tokens[j - 1].t1 > tokens[j].t0) {
            tokens[j].t0 = tokens[j - 1].t1;
            tokens[j].t1 = std::max(tokens[j].t0, tokens[j].t1);
        }
    }
    return true;
}

// Part 2, host form (sequential mode: one long buffer whose energy envelope lives on the host): the energy "VAD" pass.  The
// batched mode runs the same arithmetic on the device (a5_thold_kernel / a5_adjust_kernel below).
static void token_level_timestamps_energy(std::vector<wdr_token_data>& tokens, const Vocab& v, const float* energy, int n_samples) {
    const int n = (int)tokens.size();
    const int hw = WDR_SAMPLE_RATE / 8;
    for (int j = 0; j < n; j++) {
        if (tokens[j].id >= v.eot) continue;
        int s0 = timestamp_to_sample(tokens[j].t0, n_samples);
        int s1 = timestamp_to_sample(tokens[j].t1, n_samples);
        const int ss0 = std::max(s0 - hw, 0);
        const int ss1 = std::min(s1 + hw, n_samples);
        const int ns = ss1 - ss0;
        float sum = 0.0f;
        for (int k = ss0; k < ss1; k++) sum += energy[k];
        const float thold = 0.5f * sum / ns;
        {
            int k = s0;
            if (energy[k] > thold && j > 0) {
                while (k > 0 && energy[k] > thold) k--;
                tokens[j].t0 = sample_to_timestamp(k);
                if (tokens[j].t0 < tokens[j - 1].t1) tokens[j].t0 = tokens[j - 1].t1;
                else s0 = k;
            } else {
                while (energy[k] < thold && k < s1) k++;
                s0 = k;
                tokens[j].t0 = sample_to_timestamp(k);
            }
        }
        {
            int k = s1;
            if (energy[k] > thold) {
                while (k < n_samples - 1 && energy[k] > thold) k++;
                tokens[j].t1 = sample_to_timestamp(k);
                if (j < ns - 1 && j + 1 < n && tokens[j].t1 > tokens[j + 1].t0) tokens[j].t1 = tokens[j + 1].t0;
                else s1 = k;
            } else {
                while (energy[k] < thold && k > s0) k--;
                s1 = k;
                tokens[j].t1 = sample_to_timestamp(k);
            }
        }
    }
}

static void token_level_timestamps(std::vector<wdr_token_data>& tokens, int64_t t0, int64_t t1, const Vocab& v, const float* energy,
                                   int n_samples, float thold_pt, float thold_ptsum, int64_t* st3) {
    if (token_level_timestamps_part1(tokens, t0, t1, v, n_samples, thold_pt, thold_ptsum, st3)) token_level_timestamps_energy(tokens, v, energy, n_samples);
}

// ---- the energy pass of A.5 on the device (batched mode): the envelope never leaves HBM ----
// One entry per token of a window's segment: t0 / t1 after part 1 (centiseconds), text != 0 for ids below EOT.
struct A5Token { int64_t t0, t1; };
__device__ __forceinline__ int a5_ts_to_sample(int64_t t, int n_samples) {
    int64_t s = (t * WDR_SAMPLE_RATE) / 100;
    if (s > n_samples - 1) s = n_samples - 1;
    if (s < 0) s = 0;
    return (int)s;
}
// thold[j] = 0.5f * (energy[ss0] + ... + energy[ss1 - 1]) / ns, the sum taken sequentially in fp32 exactly as the scalar loop does.
// One WARP per token: 32 consecutive samples per coalesced load (the next chunk is requested before the current one is added), added
// in index order through shuffles — the same chain of fp32 adds.
__global__ void __launch_bounds__(256)
a5_thold_kernel(const A5Token* __restrict__ tok, const uint8_t* __restrict__ is_text, const int32_t* __restrict__ n_tok, const int32_t* __restrict__ n_valid,
                const float* __restrict__ energy, int64_t energy_stride, float* __restrict__ thold) {
    const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = n_tok[b], n_samples = n_valid[b];
    const float* e = energy + (int64_t)b * energy_stride;
    for (int j = blockIdx.x * (blockDim.x >> 5) + warp; j < n; j += gridDim.x * (blockDim.x >> 5)) {
        const int64_t o = (int64_t)b * kDecMaxTokens + j;
        if (!is_text[o]) continue;
        const int hw = WDR_SAMPLE_RATE / 8;
        const int s0 = a5_ts_to_sample(tok[o].t0, n_samples), s1 = a5_ts_to_sample(tok[o].t1, n_samples);
        const int ss0 = max(s0 - hw, 0), ss1 = min(s1 + hw, n_samples), ns = ss1 - ss0;
        float sum = 0.0f;
        float x = (ss0 + lane < ss1) ? e[ss0 + lane] : 0.0f;
        for (int k0 = ss0; k0 < ss1; k0 += 32) {
            const int kn = k0 + 32 + lane;
            const float xn = kn < ss1 ? e[kn] : 0.0f;  // in flight under the adds below
            const int cnt = min(32, ss1 - k0);
            if (cnt == 32) {
#pragma unroll
                for (int l = 0; l < 32; l++) sum += __shfl_sync(0xffffffffu, x, l);  // every lane carries the same running sum
            } else {
                for (int l = 0; l < cnt; l++) sum += __shfl_sync(0xffffffffu, x, l);
            }
            x = xn;
        }
        if (lane == 0) thold[o] = __fdiv_rn(0.5f * sum, (float)ns);
    }
}

// The expand / contract scans of every text token, tokens in order (token j reads token j-1's new t1): one CTA per window.  A scan
// `while (k > stop && cond(e[k])) k--` (or the upward form) may run over hundreds of thousands of samples when a stretch of speech
// stays above its local threshold, so the CTA examines kA5Step samples per step — 16 independent loads per thread, one block-wide
// minimum — and returns the first sample (in scan order) that ends the scalar loop: the same k the scalar code stops at.
constexpr int kA5Threads = 256, kA5PerThread = 16, kA5Step = kA5Threads * kA5PerThread;
template <bool DOWN>
__device__ __forceinline__ int a5_scan(const float* __restrict__ e, int k, int stop_at, float thold, bool while_above, int* s_first) {
    const int tid = threadIdx.x;
    while (true) {
        if (tid == 0) *s_first = INT_MAX;
        __syncthreads();
        float x[kA5PerThread];
        const int p0 = tid * kA5PerThread;  // position in scan order of this thread's first sample
#pragma unroll
        for (int j = 0; j < kA5PerThread; j++) {
            const int i = DOWN ? k - (p0 + j) : k + (p0 + j);
            const bool inside = DOWN ? i > stop_at : i < stop_at;
            x[j] = inside ? e[i] : (while_above ? -INFINITY : INFINITY);  // outside the range the scalar loop has ended: cond is false
        }
        int first = INT_MAX;
#pragma unroll
        for (int j = kA5PerThread - 1; j >= 0; j--) {
            const bool ends = while_above ? !(x[j] > thold) : !(x[j] < thold);
            if (ends) first = p0 + j;
        }
        first = __reduce_min_sync(0xffffffffu, first);
        if ((tid & 31) == 0 && first != INT_MAX) atomicMin(s_first, first);
        __syncthreads();
        const int f = *s_first;
        __syncthreads();  // everyone has read it before the next step resets it
        if (f != INT_MAX) return DOWN ? k - f : k + f;
        k = DOWN ? k - kA5Step : k + kA5Step;
    }
}
__global__ void __launch_bounds__(kA5Threads)
a5_adjust_kernel(A5Token* __restrict__ tok, const uint8_t* __restrict__ is_text, const int32_t* __restrict__ n_tok, const int32_t* __restrict__ n_valid,
                 const float* __restrict__ energy, int64_t energy_stride, const float* __restrict__ thold_all) {
    __shared__ int s_first;
    const int b = blockIdx.x;
    const int n = n_tok[b], n_samples = n_valid[b];
    if (n < 2 || n_samples <= 0) return;
    const float* e = energy + (int64_t)b * energy_stride;
    A5Token* t = tok + (int64_t)b * kDecMaxTokens;
    const int hw = WDR_SAMPLE_RATE / 8;
    for (int j = 0; j < n; j++) {  // block-uniform control flow throughout
        if (!is_text[(int64_t)b * kDecMaxTokens + j]) continue;
        int64_t tj0 = t[j].t0, tj1 = t[j].t1;
        int s0 = a5_ts_to_sample(tj0, n_samples), s1 = a5_ts_to_sample(tj1, n_samples);
        const int ns = min(s1 + hw, n_samples) - max(s0 - hw, 0);
        const float thold = thold_all[(int64_t)b * kDecMaxTokens + j];
        {
            int k = s0;
            if (e[k] > thold && j > 0) {
                k = a5_scan<true>(e, k, 0, thold, true, &s_first);              // while (k > 0 && e[k] > thold) k--
                tj0 = (100ll * k) / WDR_SAMPLE_RATE;
                const int64_t prev_t1 = t[j - 1].t1;
                if (tj0 < prev_t1) tj0 = prev_t1;
                else s0 = k;
            } else {
                k = a5_scan<false>(e, k, s1, thold, false, &s_first);           // while (e[k] < thold && k < s1) k++
                s0 = k;
                tj0 = (100ll * k) / WDR_SAMPLE_RATE;
            }
        }
        {
            int k = s1;
            if (e[k] > thold) {
                k = a5_scan<false>(e, k, n_samples - 1, thold, true, &s_first);  // while (k < n_samples - 1 && e[k] > thold) k++
                tj1 = (100ll * k) / WDR_SAMPLE_RATE;
                if (j < ns - 1 && j + 1 < n && tj1 > t[j + 1].t0) tj1 = t[j + 1].t0;
                else s1 = k;
            } else {
                k = a5_scan<true>(e, k, s0, thold, false, &s_first);            // while (e[k] < thold && k > s0) k--
                s1 = k;
                tj1 = (100ll * k) / WDR_SAMPLE_RATE;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) { t[j].t0 = tj0; t[j].t1 = tj1; }
        __syncthreads();  // token j + 1 reads t[j].t1
    }
}

// the context's vocabulary: special-token ids from n_vocab; text-token strings (and the " " token suppress_blank masks) from the
// checkpoint when the context was loaded from a ggml file
static Vocab ctx_vocab(const wdr_context* ctx) {
    Vocab v = make_vocab(ctx->arch.n_vocab);
    if (!ctx->file_tokens.empty()) {
        v.file_tokens = &ctx->file_tokens;
        for (int i = 0; i < (int)ctx->file_tokens.size() && i < v.eot; i++)
            if (ctx->file_tokens[i] == " ") { v.space = i; break; }
    }
    return v;
}

static SampleParams make_sample_params(const Vocab& v, const wdr_full_params& p) {
    SampleParams sp;
    sp.n_vocab = v.n_vocab; sp.eot = v.eot; sp.sot = v.sot; sp.translate = v.translate; sp.transcribe = v.transcribe; sp.solm = v.solm;
    sp.prev = v.prev; sp.nosp = v.nosp; sp.not_ = v.not_; sp.beg = v.beg; sp.lang0 = v.lang0; sp.n_langs = v.n_langs; sp.space = v.space;
    sp.suppress_blank = p.suppress_blank; sp.no_timestamps = p.no_timestamps; sp.single_segment = p.single_segment;
    sp.delta_min = kDeltaMin;
    sp.n_max = WDR_TEXT_CTX / 2 - 4;
    sp.initial_tid0 = p.max_initial_ts > 0.0f ? (int)roundf(p.max_initial_ts / (30.0f / 1500.0f)) : -1;
    return sp;
}

static int validate_params(const wdr_context* ctx, const wdr_full_params& p, int* lang_id) {
    if (p.strategy == WDR_SAMPLING_BEAM_SEARCH && p.beam_size > kBeamMax) { set_error("beam_size %d exceeds the supported maximum of %d", p.beam_size, kBeamMax); return WDR_ERR_UNSUPPORTED; }
    // the crate forwards advanced.temperature (src/transcribe.rs:58-68): the ladder may start anywhere in [0, 1]
    if (!(p.temperature >= 0.0f) || p.temperature > 1.0f) { set_error("temperature %g outside [0, 1]", p.temperature); return WDR_ERR_INVALID; }
    if (p.temperature_inc < 0.0f) { set_error("temperature_inc must be >= 0"); return WDR_ERR_INVALID; }
    if ((p.temperature_inc > 0.0f || p.temperature > 0.0f) && p.greedy_best_of > kBeamMax) { set_error("best_of %d exceeds the supported maximum of %d", p.greedy_best_of, kBeamMax); return WDR_ERR_UNSUPPORTED; }
    if (!p.single_segment) { set_error("single_segment = 0 is not implemented (the crate always sets it, src/transcribe.rs:46)"); return WDR_ERR_UNSUPPORTED; }
    if (p.no_timestamps) { set_error("no_timestamps = 1 is not implemented"); return WDR_ERR_UNSUPPORTED; }
    // parameters that are not implemented are refused, never silently ignored — whatever the language mode
    if (p.offset_ms || p.duration_ms || p.max_len || p.max_tokens || p.audio_ctx || p.suppress_nst) { set_error("offset/duration/max_len/max_tokens/audio_ctx/suppress_nst are not implemented"); return WDR_ERR_UNSUPPORTED; }
    if (p.translate && !ctx->arch.multilingual) { set_error("translate needs a multilingual model"); return WDR_ERR_INVALID; }
    if (p.detect_language || (p.language && strcmp(p.language, "auto") == 0)) {  // whisper_lang_auto_detect: decided per buffer on the device
        if (!ctx->arch.multilingual) { set_error("language auto-detection needs a multilingual model (whisper_lang_auto_detect fails the same way)"); return WDR_ERR_INVALID; }
        *lang_id = -1;
        return WDR_OK;
    }
    int id = lang_id_from_str(p.language ? p.language : "en");
    if (id < 0) { set_error("unknown language '%s'", p.language); return WDR_ERR_INVALID; }
    *lang_id = id;
    return WDR_OK;
}

int tokenize_for_context(const wdr_context* ctx, const char* text, std::vector<int32_t>& out);  // tokenizer.cu

// whisper_full: `initial_prompt` is tokenised first and then IS the prompt-token list (it replaces params.prompt_tokens)
static void resolve_initial_prompt(const wdr_context* ctx, wdr_full_params& p, std::vector<int32_t>& store) {
    if (!ctx || !p.initial_prompt || !p.initial_prompt[0]) return;
    tokenize_for_context(ctx, p.initial_prompt, store);
    p.prompt_tokens = store.data();
    p.prompt_n_tokens = (int)store.size();
    p.initial_prompt = nullptr;
}

template <typename T>
static int grow_dev(T** p, size_t* cap, size_t need) {
    if (need <= *cap) return WDR_OK;
    if (*p) cudaFree(*p);
    *p = nullptr; *cap = 0;
    WDR_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(p), need * sizeof(T)));
    *cap = need;
    return WDR_OK;
}
template <typename T>
static int grow_pinned(T** p, size_t* cap, size_t need) {
    if (need <= *cap) return WDR_OK;
    if (*p) cudaFreeHost(*p);
    *p = nullptr; *cap = 0;
    WDR_CUDA_TRY(cudaMallocHost(reinterpret_cast<void**>(p), need * sizeof(T)));
    *cap = need;
    return WDR_OK;
}

// ---------------------------------------------------------------------------------------------------
// Beam search and the temperature ladder (whisper_full with WHISPER_SAMPLING_BEAM_SEARCH — the crate's default strategy,
// reference src/transcribe.rs:22, 29-32 — at one temperature of `for t = temperature; t < 1 + 1e-6; t += temperature_inc`).
// Rows of the decode batch = the windows of this pass x Kd decoders (T = 0: Kd = beam_size; T > 0: Kd = max(1, greedy.best_of),
// which whisper_full_default_params leaves at -1 for the beam strategy, so one decoder).  Per iteration the device scores every live
// row (whisper_process_logits on logits / T + the Kc = beam_size best tokens, dec_topk_kernel); the host runs whisper.cpp's candidate
// logic per window — Kc candidates per live decoder, stable sort by cumulative log-probability, each live decoder takes the next
// candidate that is not a duplicate of the previous one (after the first iteration), then the per-decoder bookkeeping of the greedy
// loop (timestamp pairing, seek_delta, completion, failure) — and sends back parents + tokens; the self cache is never copied:
// beam_anc_kernel re-threads an ancestry table that the self-attention kernel reads through.  Ranking (whisper_sequence_score):
// decoders that failed are skipped, a decoder whose last 32 kept tokens have entropy < entropy_thold (result_len > 32) fails,
// the best score (sum of log-probabilities / length, or the length_penalty form) wins, first maximum, decoder 0 if none.
// ---------------------------------------------------------------------------------------------------
// whisper_sample_token(best = false): std::discrete_distribution over probs = expf(logprobs) drawn with the decoder's std::mt19937 —
// the very libstdc++ objects whisper.cpp uses, so a draw is bit-identical given identical log-probabilities.
static int sample_discrete(const float* logprobs, int n, std::mt19937& rng, std::vector<float>& probs) {
    probs.resize(n);
    for (int i = 0; i < n; i++) probs[i] = logprobs[i] == -INFINITY ? 0.0f : expf(logprobs[i]);
    std::discrete_distribution<> dist(probs.begin(), probs.end());
    return dist(rng);
}

struct HostBeam {
    std::vector<wdr_token_data> tokens;
    double sum_all = 0.0;
    int has_ts = 0, seek_delta = 100 * 30, result_len = 0;
    bool failed = false, completed = false;
};

// entropy of the ids of the last 32 kept tokens (whisper_sequence_score: std::map order = ascending id)
static double sequence_entropy(const wdr_token_data* t, int result_len) {
    std::map<int32_t, int> counts;
    int cnt = 0;
    for (int i = std::max(0, result_len - 32); i < result_len; i++) { counts[t[i].id]++; cnt++; }
    double e = 0.0;
    for (auto& kv : counts) { const double pr = (double)kv.second / cnt; e -= pr * log(pr); }
    return e;
}

static double sequence_score(const wdr_token_data* t, int result_len, float length_penalty, double* avg) {
    double sum = 0.0;
    for (int j = 0; j < result_len; j++) sum += t[j].plog;
    *avg = sum / result_len;
    double penalty = result_len;
    if (length_penalty > 0.0f) penalty = pow((5.0 + penalty) / 6.0, (double)length_penalty);
    return sum / penalty;
}

// wins: the windows of this pass (indices into win / toks_out / seq_win).  On return win[w] / toks_out[w] hold the best decoder of
// every window w of the pass and dec_failed[w] says whether that decoder counts as failed for the fallback test.
static int beam_decode(wdr_context* ctx, wdr_state* st, DecoderWorkspace& ws, const wdr_full_params& p, const SampleParams& sp, const Vocab& v,
                       const std::vector<int>& wins, int Kd, int Kc, float temperature, int n_prompt, const std::vector<int32_t>& seq_win /* [B][448] */,
                       std::vector<DecWinState>& win, std::vector<wdr_token_data>& toks_out /* [B][224] */, std::vector<char>& dec_failed, int* steps_out) {
    cudaStream_t s = st->stream;
    const int nW = (int)wins.size(), K = Kd;
    const int R = nW * K, n_max = sp.n_max;
    WDR_REQUIRE(R >= 1 && R <= kDecMaxBatch && R <= ws.cap_B && Kc >= 0 && Kc <= kBeamMax, "beam_decode: bad batch");
    int rc;
    // Kc == 0: the greedy strategy above temperature 0 — every decoder draws its next token from its own distribution
    // (whisper_sample_token, best = false) with its own generator; decoder j of a window starts from std::mt19937(j).
    const bool sampling = Kc == 0;
    FullScratch& fs = st->full;
    std::vector<std::mt19937> rngs;
    if (sampling) {
        WDR_REQUIRE(temperature > 0.0f, "sampling needs a temperature above 0");
        if ((rc = grow_dev(&fs.lp_dev, &fs.lp_dev_cap, (size_t)R * ws.ldv)) != WDR_OK) return rc;
        if ((rc = grow_pinned(&fs.lp_host, &fs.lp_host_cap, (size_t)R * ws.ldv)) != WDR_OK) return rc;
        for (int r = 0; r < R; r++) rngs.emplace_back((uint32_t)(r % K));
    }
    std::vector<int32_t> seq((size_t)R * kDecSeqCap);
    std::vector<int32_t> rowwin(kDecMaxBatch, 0);
    for (int r = 0; r < R; r++) {
        memcpy(&seq[(size_t)r * kDecSeqCap], &seq_win[(size_t)wins[r / K] * kDecSeqCap], sizeof(int32_t) * kDecSeqCap);
        rowwin[r] = wins[r / K];
    }
    std::vector<int32_t> anc((size_t)kDecSeqCap * kDecMaxBatch);
    for (int b = 0; b < kDecMaxBatch; b++)  // row-major [row][448]: every position of a fresh row lives in the row itself
        for (int t = 0; t < kDecSeqCap; t++) anc[(size_t)b * kDecSeqCap + t] = b;
    std::vector<int32_t> limit(kDecMaxBatch, 0);
    for (int r = 0; r < R; r++) limit[r] = win[wins[r / K]].completed ? 0 : 1 << 30;
    ws.beam_anc_cur = ws.beam_anc[0];
    ws.beam_K = K;
    WDR_CUDA_TRY(cudaMemcpyAsync(ws.seq, seq.data(), sizeof(int32_t) * seq.size(), cudaMemcpyHostToDevice, s));
    WDR_CUDA_TRY(cudaMemcpyAsync(ws.beam_anc[0], anc.data(), sizeof(int32_t) * anc.size(), cudaMemcpyHostToDevice, s));
    WDR_CUDA_TRY(cudaMemcpyAsync(ws.beam_limit, limit.data(), sizeof(int32_t) * kDecMaxBatch, cudaMemcpyHostToDevice, s));
    WDR_CUDA_TRY(cudaMemcpyAsync(ws.beam_rowwin, rowwin.data(), sizeof(int32_t) * kDecMaxBatch, cudaMemcpyHostToDevice, s));
    for (int i = 0; i < n_prompt; i++)
        if ((rc = decoder_step(ctx, ws, R, i, i == n_prompt - 1, DEC_MODE_BEAM, s, &st->prof)) != WDR_OK) return rc;
    std::vector<std::vector<HostBeam>> beams(nW, std::vector<HostBeam>(K));
    std::vector<BeamRow> rows(R);
    std::vector<BeamCand> cands((size_t)R * kBeamMax);
    std::vector<float> nosp(R, 0.0f);
    std::vector<int32_t> parent(R), next_tok(R);
    int steps = 0;
    struct Cand { int k, r; double sum; };
    for (int i = 0; i < n_max; i++) {
        for (int r = 0; r < R; r++) {
            const HostBeam& hb = beams[r / K][r % K];
            BeamRow br;
            br.active = !win[wins[r / K]].completed && !hb.completed && !hb.failed;
            br.n_cur = (int)hb.tokens.size();
            br.last_id = br.n_cur > 0 ? hb.tokens[br.n_cur - 1].id : 0;
            br.penult_id = br.n_cur > 1 ? hb.tokens[br.n_cur - 2].id : 0;
            br.has_ts = hb.has_ts;
            br.seek_delta = hb.seek_delta;
            rows[r] = br;
        }
        WDR_CUDA_TRY(cudaMemcpyAsync(ws.beam_rows, rows.data(), sizeof(BeamRow) * R, cudaMemcpyHostToDevice, s));
        if ((rc = decoder_topk(ctx, ws, R, sp, sampling ? 1 : Kc, temperature, sampling ? fs.lp_dev : nullptr, s, &st->prof)) != WDR_OK) return rc;
        WDR_CUDA_TRY(cudaMemcpyAsync(cands.data(), ws.beam_cands, sizeof(BeamCand) * (size_t)R * kBeamMax, cudaMemcpyDeviceToHost, s));
        if (sampling)
            for (int r = 0; r < R; r++)
                if (rows[r].active)
                    WDR_CUDA_TRY(cudaMemcpyAsync(fs.lp_host + (size_t)r * ws.ldv, fs.lp_dev + (size_t)r * ws.ldv, sizeof(float) * sp.n_vocab, cudaMemcpyDeviceToHost, s));
        if (i == 0) WDR_CUDA_TRY(cudaMemcpyAsync(nosp.data(), ws.beam_nosp, sizeof(float) * R, cudaMemcpyDeviceToHost, s));
        WDR_CUDA_TRY(cudaStreamSynchronize(s));
        steps = i + 1;
        if (i == 0)
            for (int w = 0; w < nW; w++) win[wins[w]].no_speech_prob = nosp[(size_t)w * K];
        bool any_live = false;
        for (int r = 0; r < R; r++) { parent[r] = r; next_tok[r] = v.eot; }
        if (sampling) {  // draw on the host; slot 0 of the row's candidates becomes the drawn token
            parallel_for(R, [&](int r) {
                if (!rows[r].active) return;
                std::vector<float> probs;
                const float* lp = fs.lp_host + (size_t)r * ws.ldv;
                const int id = sample_discrete(lp, sp.n_vocab, rngs[r], probs);
                const BeamCand raw = cands[(size_t)r * kBeamMax + 1];
                BeamCand c;
                c.id = id; c.tid = raw.tid; c.p = probs[id]; c.plog = lp[id]; c.pt = raw.pt; c.ptsum = raw.ptsum;
                if (id >= v.beg) { c.tid = id; c.pt = c.p; }
                cands[(size_t)r * kBeamMax] = c;
            });
        }
        for (int w = 0; w < nW; w++) {
            if (win[wins[w]].completed) continue;
            std::vector<HostBeam>& bw = beams[w];
            std::vector<Cand> cl;
            for (int k = 0; k < K; k++) {
                if (bw[k].completed || bw[k].failed) continue;
                for (int r = 0; r < (sampling ? 1 : Kc); r++) {
                    const BeamCand& c = cands[((size_t)w * K + k) * kBeamMax + r];
                    if (c.id >= 0) cl.push_back({k, r, bw[k].sum_all + (double)c.plog});
                }
            }
            if (cl.empty()) continue;
            if (!sampling) std::stable_sort(cl.begin(), cl.end(), [](const Cand& a, const Cand& b) { return a.sum > b.sum; });
            auto cand_tok = [&](const Cand& c) -> const BeamCand& { return cands[((size_t)w * K + c.k) * kBeamMax + c.r]; };
            auto same_seq = [&](const Cand& a, const Cand& b) {
                if (cand_tok(a).id != cand_tok(b).id) return false;
                const auto &ta = bw[a.k].tokens, &tb = bw[b.k].tokens;
                if (ta.size() != tb.size()) return false;
                for (size_t j = 0; j < ta.size(); j++)
                    if (ta[j].id != tb[j].id) return false;
                return true;
            };
            std::vector<HostBeam> nb = bw;
            size_t cur_c = 0;
            for (int k = 0; k < K; k++) {
                if (bw[k].completed || bw[k].failed) continue;
                if (cur_c >= cl.size()) cur_c = 0;
                const Cand cur = cl[cur_c++];
                while (!sampling && cl.size() > cur_c && i > 0 && same_seq(cl[cur_c], cur)) ++cur_c;
                const BeamCand& c = cand_tok(cur);
                HostBeam h = bw[cur.k];
                wdr_token_data td;
                memset(&td, 0, sizeof(td));
                td.id = c.id; td.tid = c.tid; td.p = c.p; td.plog = c.plog; td.pt = c.pt; td.ptsum = c.ptsum; td.t0 = -1; td.t1 = -1; td.t_dtw = -1;
                h.tokens.push_back(td);
                h.sum_all = cur.sum;
                // ---- per-decoder bookkeeping (the greedy loop's rules) ----
                const int seek = win[wins[w]].seek, seek_end = win[wins[w]].seek_end;
                bool done = false;
                if (td.id > v.beg) {
                    const int sd_new = 2 * (td.id - v.beg);
                    if (h.has_ts && h.seek_delta > sd_new && h.result_len < i) { h.failed = true; done = true; }
                    else { h.seek_delta = sd_new; h.result_len = i + 1; h.has_ts = 1; }
                }
                if (!done && (td.id == v.eot || (h.has_ts && seek + h.seek_delta + sp.delta_min >= seek_end))) {
                    bool fail = false;
                    if (h.result_len == 0 && !sp.no_timestamps) {
                        if (seek + h.seek_delta + sp.delta_min >= seek_end) h.result_len = i + 1;
                        else fail = true;
                    }
                    if (fail) h.failed = true;
                    else {
                        if (sp.single_segment || sp.no_timestamps) { h.result_len = i + 1; h.seek_delta = 100 * 30; }
                        h.completed = true;
                    }
                    done = true;
                }
                if (!done && i == n_max - 1 && (h.result_len == 0 || h.seek_delta < 100 * 30 / 2)) { h.failed = true; done = true; }
                nb[k] = h;
                parent[(size_t)w * K + k] = w * K + cur.k;
                next_tok[(size_t)w * K + k] = td.id;
                if (!done) any_live = true;
            }
            bw.swap(nb);
        }
        if (!any_live || i == n_max - 1) break;
        if (p.abort_callback && (i & 7) == 7 && p.abort_callback(p.abort_callback_user_data)) { set_error("aborted by callback"); return WDR_ERR_ABORTED; }
        // ---- next step: re-thread the ancestry, feed the chosen tokens, skip dead rows ----
        const int pos_last = n_prompt - 1 + i;
        for (int r = 0; r < R; r++) {
            const HostBeam& hb = beams[r / K][r % K];
            limit[r] = (win[wins[r / K]].completed || hb.completed || hb.failed) ? 0 : 1 << 30;
        }
        WDR_CUDA_TRY(cudaMemcpyAsync(ws.beam_parent, parent.data(), sizeof(int32_t) * R, cudaMemcpyHostToDevice, s));
        WDR_CUDA_TRY(cudaMemcpyAsync(ws.beam_limit, limit.data(), sizeof(int32_t) * R, cudaMemcpyHostToDevice, s));
        WDR_CUDA_TRY(cudaMemcpy2DAsync(ws.seq + pos_last + 1, sizeof(int32_t) * kDecSeqCap, next_tok.data(), sizeof(int32_t), sizeof(int32_t), R, cudaMemcpyHostToDevice, s));
        if ((rc = decoder_beam_reorder(ws, R, pos_last, s)) != WDR_OK) return rc;
        if ((rc = decoder_step(ctx, ws, R, pos_last + 1, true, DEC_MODE_BEAM, s, &st->prof)) != WDR_OK) return rc;
        WDR_CUDA_TRY(cudaStreamSynchronize(s));  // parent / limit / next_tok are reused by the next iteration
    }
    // ---- rank the decoders of every window (whisper_sequence_score + the entropy test) ----
    for (int w = 0; w < nW; w++) {
        DecWinState& ww = win[wins[w]];
        if (ww.completed) continue;  // was never decoded (too short)
        int best = 0;
        double best_score = -INFINITY;
        std::vector<char> failed_k(K, 0);
        for (int k = 0; k < K; k++) {
            const HostBeam& h = beams[w][k];
            failed_k[k] = h.failed;
            if (h.failed) continue;
            if (h.result_len <= 0) { failed_k[k] = 1; continue; }  // nothing kept: whisper_sequence_score returns early and the stale score never wins
            double avg;
            const double score = sequence_score(h.tokens.data(), h.result_len, p.length_penalty, &avg);
            if (h.result_len > 32 && sequence_entropy(h.tokens.data(), h.result_len) < (double)p.entropy_thold) { failed_k[k] = 1; continue; }
            if (best_score < score) { best = k; best_score = score; }
        }
        const HostBeam& h = beams[w][best];
        ww.n_cur = (int)h.tokens.size(); ww.has_ts = h.has_ts; ww.seek_delta = h.seek_delta; ww.result_len = h.result_len;
        ww.failed = h.failed ? 1 : 0; ww.completed = h.completed ? 1 : 0;
        dec_failed[wins[w]] = failed_k[best];
        for (size_t j = 0; j < h.tokens.size() && j < (size_t)kDecMaxTokens; j++) toks_out[(size_t)wins[w] * kDecMaxTokens + j] = h.tokens[j];
    }
    *steps_out = steps;
    return WDR_OK;
}

// Sequential mode (whisper_full's own seek loop over a buffer longer than 30 s): one window of the long buffer.
struct SeqWindow {
    const float* mel_dev;         // raw log-mel of the WHOLE buffer [n_mel][n_len] (device)
    int n_len;
    const float* max_dev;         // its global maximum (whisper.cpp normalises with the buffer-global max)
    int seek, seek_end;           // window start / end of audio, mel frames (= centiseconds)
    const float* energy_host;     // get_signal_energy of the whole buffer (host)
    int n_samples;
    int64_t* st3;                 // t_beg, t_last, tid_last carried across the call's segments
    const std::vector<int32_t>* prompt_past;
    std::vector<int32_t>* result_ids;  // out: tokens_cur[0 .. result_len) of the window unless it is a no-speech window (what whisper_full appends to prompt_past)
};

// One group of B <= 128 windows whose PCM is already on the device (In = int16_t or float).  sw != nullptr: B == 1 and the
// window comes from a long buffer's mel at frame sw->seek instead of from pcm_dev.
template <typename In>
static int full_group(wdr_context* ctx, wdr_state* st, const wdr_full_params& p, int lang_id, const In* pcm_dev, int64_t chunk_stride,
                      const int32_t* n_valid_host, int chunk0, int B, const SeqWindow* sw = nullptr) {
    FullScratch& fs = st->full;
    DecoderWorkspace& ws = st->dec;
    const WhisperArch& a = ctx->arch;
    const Vocab v = ctx_vocab(ctx);
    cudaStream_t s = st->stream;
    int rc;
    const int beam_K = (p.strategy == WDR_SAMPLING_BEAM_SEARCH && p.beam_size > 1) ? p.beam_size : 1;  // rows per window of the decode batch
    WDR_REQUIRE(B <= kDecMaxWindows && B * beam_K <= kDecMaxRows, "windows x beams exceeds the decode batch (128 windows, 640 rows)");
    const int fb_Kd = std::max(1, p.greedy_best_of);  // decoders per window of a fallback pass (temperature > 0)
    const int fb_rows = (p.temperature_inc > 0.0f || p.temperature > 0.0f) ? std::min(B, kDecMaxRows / fb_Kd) * fb_Kd : 0;
    if ((rc = ws.reserve(ctx, B, std::max(std::max(B * beam_K, B), fb_rows))) != WDR_OK) return rc;
    // ---- n_valid on the device ----
    if ((rc = grow_dev(&fs.nvalid_dev, &fs.nvalid_cap, (size_t)B)) != WDR_OK) return rc;
    std::vector<int32_t> nv(B);
    for (int b = 0; b < B; b++) nv[b] = n_valid_host ? n_valid_host[chunk0 + b] : WDR_CHUNK_SAMPLES;
    WDR_CUDA_TRY(cudaMemcpyAsync(fs.nvalid_dev, nv.data(), sizeof(int32_t) * B, cudaMemcpyHostToDevice, s));
    for (auto& e : fs.ev_phase)
        if (!e) WDR_CUDA_TRY(cudaEventCreate(&e));
    WDR_CUDA_TRY(cudaEventRecord(fs.ev_phase[0], s));
    // ---- mel + encoder -> bf16 hidden states ----
    if (sw) {
        WDR_REQUIRE(B == 1, "sequential mode decodes one window at a time");
        if ((rc = st->enc.reserve(ctx->arch, 1)) != WDR_OK) return rc;
        if ((rc = encoder_forward(ctx, st->enc, sw->mel_dev, sw->n_len, sw->seek, sw->max_dev, 0, 1, nullptr, ws.enc_bf16, s, &st->prof)) != WDR_OK) return rc;
    } else if ((rc = encode_chunks<In>(ctx, st->enc, pcm_dev, chunk_stride, fs.nvalid_dev, B, nullptr, ws.enc_bf16, s, &st->prof)) != WDR_OK) return rc;
    // ---- energy for the token-timestamp heuristic (D2H overlaps the decode) ----
    // The envelope stays in HBM: the energy pass of A.5 runs on the device (a5_thold_kernel / a5_adjust_kernel) after the decode.
    if (p.token_timestamps && !sw) {
        if ((rc = grow_dev(&fs.energy_dev, &fs.energy_cap, (size_t)B * WDR_CHUNK_SAMPLES)) != WDR_OK) return rc;
        ProfScope ps(&st->prof, KC_OTHER, s);
        energy_kernel<In><<<dim3(256, B), 256, 0, s>>>(pcm_dev, chunk_stride, fs.nvalid_dev, WDR_CHUNK_SAMPLES, 32, fs.energy_dev, WDR_CHUNK_SAMPLES);
        WDR_LAUNCH_CHECK();
        WDR_CUDA_TRY(cudaEventRecord(fs.ev_energy, s));  // the A.5 kernels run on the side stream, under the DTW pass
    }
    WDR_CUDA_TRY(cudaEventRecord(fs.ev_phase[1], s));
    WDR_CUDA_TRY(cudaMemsetAsync(ws.cross_stats, 0, sizeof(unsigned long long) * 2, s));
    if ((rc = decoder_cross_kv(ctx, ws, B, s, &st->prof)) != WDR_OK) return rc;
    WDR_CUDA_TRY(cudaEventRecord(fs.ev_phase[2], s));
    // ---- language: given, or whisper_lang_auto_detect per buffer: decode [SOT] at position 0, arg-max over the language tokens ----
    std::vector<int> lang(B, lang_id);
    if (lang_id < 0) {
        std::vector<int32_t> sq((size_t)B * kDecSeqCap, v.eot);
        for (int b = 0; b < B; b++) sq[(size_t)b * kDecSeqCap] = v.sot;
        WDR_CUDA_TRY(cudaMemcpyAsync(ws.seq, sq.data(), sizeof(int32_t) * sq.size(), cudaMemcpyHostToDevice, s));
        if ((rc = decoder_step(ctx, ws, B, 0, true, DEC_MODE_FORCED, s, &st->prof)) != WDR_OK) return rc;
        const int n_l = v.translate - v.lang0;  // 99 languages, 100 for large-v3
        std::vector<float> lg((size_t)B * n_l);
        WDR_CUDA_TRY(cudaMemcpy2DAsync(lg.data(), sizeof(float) * n_l, ws.logits + v.lang0, sizeof(float) * ws.ldv, sizeof(float) * n_l, B, cudaMemcpyDeviceToHost, s));
        WDR_CUDA_TRY(cudaStreamSynchronize(s));  // also fences the pageable staging vector sq
        for (int b = 0; b < B; b++) {
            int best = 0;
            for (int i = 1; i < n_l; i++)
                if (lg[(size_t)b * n_l + i] > lg[(size_t)b * n_l + best]) best = i;  // softmax over the language tokens is monotonic: same arg-max
            lang[b] = best;
        }
        if (chunk0 == 0 || sw) st->lang_id = lang[0];
    }
    if (p.detect_language) {  // whisper_full returns right after the detection
        for (int b = 0; b < B; b++) {
            ChunkInfo ci;
            memset(&ci, 0, sizeof(ci));
            ci.lang_id = lang[b];
            st->chunk_info.push_back(ci);
        }
        return WDR_OK;
    }
    // ---- prompt + decoder state ----
    // whisper_full: [PREV] + the tail of prompt_past only while the temperature is below 0.5, then [SOT, (LANG, TASK)]
    const int32_t* past = sw ? sw->prompt_past->data() : p.prompt_tokens;
    const int n_past = sw ? (int)sw->prompt_past->size() : p.prompt_n_tokens;
    int lang_slot = -1;  // position of the language token in the prompt (per-window value)
    auto build_prompt = [&](bool with_past) {
        std::vector<int32_t> prompt;
        if (with_past && past && n_past > 0 && p.n_max_text_ctx > 0) {
            const int n_take = std::min(std::min(p.n_max_text_ctx, WDR_TEXT_CTX / 2), n_past);
            prompt.push_back(v.prev);
            for (int i = n_past - n_take; i < n_past; i++) prompt.push_back(past[i]);
        }
        prompt.push_back(v.sot);
        lang_slot = -1;
        if (v.multilingual) {
            lang_slot = (int)prompt.size();
            prompt.push_back(v.lang0 + std::max(lang[0], 0));
            prompt.push_back(p.translate ? v.translate : v.transcribe);
        }
        return prompt;
    };
    std::vector<float> temps;  // the temperature ladder
    if (p.temperature_inc > 0.0f) for (float t = p.temperature; t < 1.0f + 1e-6f; t += p.temperature_inc) temps.push_back(t);
    else temps.push_back(p.temperature);
    std::vector<int32_t> prompt = build_prompt(temps[0] < 0.5f);
    int n_prompt = (int)prompt.size();
    const int n_max = WDR_TEXT_CTX / 2 - 4;
    WDR_REQUIRE(n_prompt + n_max <= kDecSeqCap, "prompt too long");
    std::vector<int32_t> seq((size_t)B * kDecSeqCap, v.eot);
    std::vector<DecWinState> win(B);
    int n_skip = 0;
    auto fill_seq = [&]() {
        std::fill(seq.begin(), seq.end(), v.eot);
        for (int b = 0; b < B; b++) {
            for (int i = 0; i < n_prompt; i++) seq[(size_t)b * kDecSeqCap + i] = prompt[i];
            if (lang_slot >= 0) seq[(size_t)b * kDecSeqCap + lang_slot] = v.lang0 + lang[b];
        }
    };
    fill_seq();
    for (int b = 0; b < B; b++) {
        DecWinState w;
        memset(&w, 0, sizeof(w));
        w.seek_delta = 100 * 30;
        w.seek = sw ? sw->seek : 0;
        w.seek_end = sw ? sw->seek_end : 1 + (nv[b] + 200 - WDR_N_FFT) / WDR_HOP;  // mel.n_len_org
        if ((!sw && nv[b] <= 0) || w.seek_end < w.seek + kDeltaMin || w.seek + kDeltaMin >= w.seek_end) { w.completed = 1; w.seek_delta = 0; n_skip++; }  // too short: no decode
        win[b] = w;
    }
    const std::vector<DecWinState> win0 = win;  // the state every temperature starts from
    WDR_CUDA_TRY(cudaMemcpyAsync(ws.seq, seq.data(), sizeof(int32_t) * seq.size(), cudaMemcpyHostToDevice, s));
    WDR_CUDA_TRY(cudaMemcpyAsync(ws.win, win.data(), sizeof(DecWinState) * B, cudaMemcpyHostToDevice, s));
    WDR_CUDA_TRY(cudaMemsetAsync(ws.done_count, 0, sizeof(int32_t), s));
    WDR_CUDA_TRY(cudaMemsetAsync(ws.tokens, 0, sizeof(wdr_token_data) * (size_t)B * kDecMaxTokens, s));
    const SampleParams sp = make_sample_params(v, p);
    // ---- decode at the first temperature: beam search (rows = windows x beams) or greedy ----
    int steps_run = 0;
    std::vector<wdr_token_data> toks((size_t)B * kDecMaxTokens);
    memset(toks.data(), 0, sizeof(wdr_token_data) * toks.size());
    std::vector<char> dec_failed(B, 0);       // the best decoder counts as failed (bookkeeping failure or the entropy test)
    std::vector<float> final_temp(B, temps[0]);
    bool results_on_host = false;
    // One pass of the ladder above temperature 0 over the windows `list`: max(1, greedy.best_of) decoders per window (both strategies),
    // the beam strategy ranks beam_size candidates per decoder, the greedy strategy draws (whisper_sample_token, best = false).
    const int ladder_Kd = std::max(1, p.greedy_best_of);
    auto ladder_pass = [&](const std::vector<int>& list, size_t it) -> int {
        prompt = build_prompt(temps[it] < 0.5f);
        n_prompt = (int)prompt.size();
        fill_seq();
        const int per_pass = kDecMaxRows / ladder_Kd;
        for (size_t o = 0; o < list.size(); o += per_pass) {
            std::vector<int> sub(list.begin() + o, list.begin() + std::min(list.size(), o + per_pass));
            for (int b : sub) {
                win[b] = win0[b];
                final_temp[b] = temps[it];
                memset(&toks[(size_t)b * kDecMaxTokens], 0, sizeof(wdr_token_data) * kDecMaxTokens);
            }
            int steps_pass = 0;
            const int Kc = p.strategy == WDR_SAMPLING_BEAM_SEARCH ? std::max(1, p.beam_size) : 0;  // greedy strategy: every decoder draws (sampling)
            int rc2 = beam_decode(ctx, st, ws, p, sp, v, sub, ladder_Kd, Kc, temps[it], n_prompt, seq, win, toks, dec_failed, &steps_pass);
            if (rc2 != WDR_OK) return rc2;
            steps_run += steps_pass;
        }
        return WDR_OK;
    };
    if (temps[0] > 0.0f && n_skip < B) {  // the crate's advanced.temperature: the ladder starts above 0, every window is decoded the way a fallback pass is
        std::vector<int> all;
        for (int b = 0; b < B; b++)
            if (!win0[b].completed) all.push_back(b);
        if ((rc = ladder_pass(all, 0)) != WDR_OK) return rc;
        results_on_host = true;
    } else if (beam_K > 1 && n_skip < B) {
        std::vector<int> all;
        for (int b = 0; b < B; b++) all.push_back(b);
        if ((rc = beam_decode(ctx, st, ws, p, sp, v, all, beam_K, beam_K, temps[0], n_prompt, seq, win, toks, dec_failed, &steps_run)) != WDR_OK) return rc;
        results_on_host = true;
    } else if (n_skip < B) {
        for (int i = 0; i < n_prompt; i++)
            if ((rc = decoder_step(ctx, ws, B, i, i == n_prompt - 1, DEC_MODE_DECODE, s, &st->prof)) != WDR_OK) return rc;
        int32_t* done_host = fs.done_host;
        // Greedy loop.  Iteration i = sample at position n_prompt-1+i, then the decoder step at n_prompt+i.  Unless per-kernel
        // profiling is on, one iteration is a single CUDA-graph launch (the position lives in ws.pos_dev); the host looks at the
        // finished-window counter every kDoneCheck iterations.
        static const bool no_graph = getenv("WDR_NO_GRAPH") != nullptr;
        const bool use_graph = !st->prof.enabled && !no_graph;
        constexpr int kDoneCheck = 8;
        cudaGraphExec_t graph = nullptr;
        if (use_graph) {
            if ((rc = decoder_decode_graph(ctx, ws, B, sp, s, &graph)) != WDR_OK) return rc;
            const int32_t pos0 = n_prompt - 1;
            WDR_CUDA_TRY(cudaMemcpyAsync(ws.pos_dev, &pos0, sizeof(int32_t), cudaMemcpyHostToDevice, s));
            WDR_CUDA_TRY(cudaStreamSynchronize(s));  // pos0 is a stack variable
        }
        for (int i = 0; i < n_max; i++) {
            const bool last = i == n_max - 1;
            if (use_graph && !last) {
                WDR_CUDA_TRY(cudaGraphLaunch(graph, s));
                count_launch((uint64_t)ws.graph_nodes);
            } else {
                if ((rc = decoder_sample(ctx, ws, B, n_prompt - 1 + i, sp, s, &st->prof)) != WDR_OK) return rc;
            }
            steps_run = i + 1;
            if ((i % kDoneCheck) == kDoneCheck - 1 || last) {
                WDR_CUDA_TRY(cudaMemcpyAsync(done_host, ws.done_count, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
                WDR_CUDA_TRY(cudaStreamSynchronize(s));
                if (*done_host + n_skip >= B) break;
                if (p.abort_callback && p.abort_callback(p.abort_callback_user_data)) { set_error("aborted by callback"); return WDR_ERR_ABORTED; }
            }
            if (last) break;
            if (!use_graph && (rc = decoder_step(ctx, ws, B, n_prompt + i, true, DEC_MODE_DECODE, s, &st->prof)) != WDR_OK) return rc;
        }
    }
    // ---- temperature fallback (whisper_full: "was the decoding successful for the current temperature?") ----
    // A window whose best decoder failed, or whose average log-probability is below logprob_thold while it is not a no-speech
    // window, is decoded again at the next temperature; the last temperature's result stands whatever it is.  Windows are
    // independent, so only the failing ones go on, as rows of a fresh decode batch that read their cross caches in place.
    if (temps.size() > 1 && n_skip < B) {
        if (!results_on_host) {
            WDR_CUDA_TRY(cudaMemcpyAsync(toks.data(), ws.tokens, sizeof(wdr_token_data) * toks.size(), cudaMemcpyDeviceToHost, s));
            WDR_CUDA_TRY(cudaMemcpyAsync(win.data(), ws.win, sizeof(DecWinState) * B, cudaMemcpyDeviceToHost, s));
            WDR_CUDA_TRY(cudaStreamSynchronize(s));
            results_on_host = true;
            for (int b = 0; b < B; b++) {  // single greedy decoder: the entropy test of the ranking step
                const DecWinState& w = win[b];
                dec_failed[b] = w.failed || (!w.failed && w.result_len > 32 &&
                                             sequence_entropy(&toks[(size_t)b * kDecMaxTokens], w.result_len) < (double)p.entropy_thold);
            }
        }
        auto unsuccessful = [&](int b) {
            const DecWinState& w = win[b];
            if (dec_failed[b]) return true;
            if (w.result_len <= 0) return true;  // nothing kept: avg_logprobs is undefined upstream; treated as a failed decode
            double avg;
            sequence_score(&toks[(size_t)b * kDecMaxTokens], w.result_len, p.length_penalty, &avg);
            return avg < (double)p.logprob_thold && w.no_speech_prob < p.no_speech_thold;
        };
        std::vector<int> fb;
        for (int b = 0; b < B; b++)
            if (!win0[b].completed && unsuccessful(b)) fb.push_back(b);
        for (size_t it = 1; it < temps.size() && !fb.empty(); it++) {
            if ((rc = ladder_pass(fb, it)) != WDR_OK) return rc;
            if (it + 1 == temps.size()) break;
            std::vector<int> still;
            for (int b : fb)
                if (unsuccessful(b)) still.push_back(b);
            fb.swap(still);
        }
    }
    fs.last_decode_steps = steps_run;
    fs.decode_steps += steps_run;
    WDR_CUDA_TRY(cudaEventRecord(fs.ev_phase[3], s));
    // ---- results to the host ----
    static const bool dbg_time = getenv("WDR_DEBUG_TIMING") != nullptr;
    auto now_ms = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_h0 = now_ms();
    if (!results_on_host) {
        WDR_CUDA_TRY(cudaMemcpyAsync(toks.data(), ws.tokens, sizeof(wdr_token_data) * toks.size(), cudaMemcpyDeviceToHost, s));
        WDR_CUDA_TRY(cudaMemcpyAsync(win.data(), ws.win, sizeof(DecWinState) * B, cudaMemcpyDeviceToHost, s));
    }
    WDR_CUDA_TRY(cudaStreamSynchronize(s));
    const double t_h1 = now_ms();
    const double t_h2 = now_ms();

    struct Pending { int seg; int b; int n_frames; std::vector<int32_t> dtw_seq; int sot_len; };
    std::vector<Pending> pend;
    struct PostJob { int seg; int b; };
    std::vector<PostJob> post;
    // The heuristic token timestamps scan ~1e6 energy samples per window on the host, as whisper.cpp does: ~2 ms per
    // window.  Windows are independent, so the jobs run on a few host threads, and they run AFTER the DTW pass / DTW kernels have
    // been queued so that the GPU works underneath them.
    // Batched mode: part 1 of A.5 (scalar logic on the token list) runs on the host right where the segment is assembled; its energy
    // pass is queued on the device (the envelope never leaves HBM) and the adjusted t0 / t1 come back with the DTW paths.
    // Sequential mode (one long buffer, envelope on the host, timestamp state carried across windows): host code.
    std::vector<int> a5_seg(B, -1);   // window -> result segment whose tokens went to the device
    std::vector<int32_t> a5_n(B, 0);
    auto run_post = [&]() {
        parallel_for((int)post.size(), [&](int i) {
            ResultSegment& seg = st->results[post[i].seg];
            if (p.token_timestamps && sw) token_level_timestamps(seg.tokens, seg.t0, seg.t1, v, sw->energy_host, sw->n_samples, p.thold_pt, p.thold_ptsum, sw->st3);
            seg.token_text.reserve(seg.tokens.size());
            for (auto& t : seg.tokens) seg.token_text.push_back(token_text(v, t.id));
        });
        post.clear();
    };
    for (int b = 0; b < B; b++) {
        const DecWinState& w = win[b];
        ChunkInfo ci;
        ci.seek_delta = w.seek_delta; ci.failed = w.failed; ci.completed = w.completed; ci.n_sampled = w.n_cur; ci.has_ts = w.has_ts;
        ci.result_len = w.result_len; ci.seek_end = w.seek_end; ci.n_segments = 0; ci.no_speech_prob = w.no_speech_prob;
        ci.lang_id = lang[b];
        ci.temperature = final_temp[b];
        const int n_keep = w.failed ? w.n_cur : w.result_len;
        std::vector<wdr_token_data> cur(toks.begin() + (size_t)b * kDecMaxTokens, toks.begin() + (size_t)b * kDecMaxTokens + n_keep);
        double avg_logprob = -INFINITY;
        if (!w.failed && w.result_len > 0) {
            double sum = 0.0;
            for (int i = 0; i < w.result_len; i++) sum += cur[i].plog;
            avg_logprob = sum / w.result_len;
        }
        const bool is_no_speech = w.no_speech_prob > p.no_speech_thold && avg_logprob < p.logprob_thold;
        if (sw && sw->result_ids) {
            sw->result_ids->clear();
            if (!is_no_speech)
                for (int i = 0; i < w.result_len && i < (int)cur.size(); i++) sw->result_ids->push_back(cur[i].id);
        }
        if (!cur.empty() && !is_no_speech) {
            const int seek = w.seek;
            const int64_t t0 = seek + 2 * (cur.front().tid - v.beg);
            std::string text;
            for (auto& t : cur)
                if (p.print_special || t.id < v.eot) text += token_text(v, t.id);
            if (!text.empty()) {
                ResultSegment seg;
                seg.chunk = chunk0 + b;
                seg.t0 = t0;
                seg.t1 = seek + w.seek_delta;
                seg.text = text;
                seg.no_speech_prob = w.no_speech_prob;
                seg.tokens = cur;
                st->results.push_back(std::move(seg));
                post.push_back({(int)st->results.size() - 1, b});  // token strings (and, in sequential mode, token timestamps): after the DTW launches
                ci.n_segments = 1;
                if (p.token_timestamps && !sw) {
                    ResultSegment& rs = st->results.back();
                    int64_t st3[3] = {0, 0, 0};
                    if (token_level_timestamps_part1(rs.tokens, rs.t0, rs.t1, v, nv[b], p.thold_pt, p.thold_ptsum, st3)) {
                        a5_seg[b] = (int)st->results.size() - 1;
                        a5_n[b] = (int32_t)std::min<size_t>(rs.tokens.size(), kDecMaxTokens);
                    }
                }
                if (ctx->dtw_enabled) {
                    Pending pd;
                    pd.seg = (int)st->results.size() - 1;
                    pd.b = b;
                    pd.n_frames = std::min(std::min(100 * 30, w.seek_delta), w.seek_end - seek);
                    pd.dtw_seq.push_back(v.sot);
                    if (v.multilingual) pd.dtw_seq.push_back(v.lang0 + lang[b]);
                    pd.sot_len = (int)pd.dtw_seq.size();
                    pd.dtw_seq.push_back(v.not_);
                    for (auto& t : st->results[pd.seg].tokens)
                        if (t.id < v.eot) pd.dtw_seq.push_back(t.id);
                    pd.dtw_seq.push_back(v.eot);
                    pend.push_back(std::move(pd));
                }
            }
        }
        st->chunk_info.push_back(ci);
    }
    if (dbg_time) fprintf(stderr, "[wdr] results D2H+sync %.1f ms, host assembly %.1f ms\n", t_h1 - t_h0, now_ms() - t_h2);
    // ---- energy pass of A.5 on the device ----
    bool a5_any = false;
    for (int b = 0; b < B; b++) a5_any |= a5_n[b] > 0;
    if (a5_any) {
        const size_t n_tok = (size_t)B * kDecMaxTokens;
        if ((rc = grow_dev(&fs.a5_tok_dev, &fs.a5_tok_cap, n_tok * 2)) != WDR_OK) return rc;
        if ((rc = grow_pinned(&fs.a5_tok_host, &fs.a5_tok_host_cap, n_tok * 2)) != WDR_OK) return rc;
        if ((rc = grow_dev(&fs.a5_text_dev, &fs.a5_text_cap, n_tok)) != WDR_OK) return rc;
        if ((rc = grow_pinned(&fs.a5_text_host, &fs.a5_text_host_cap, n_tok)) != WDR_OK) return rc;
        if ((rc = grow_dev(&fs.a5_n_dev, &fs.a5_n_cap, (size_t)B)) != WDR_OK) return rc;
        if ((rc = grow_pinned(&fs.a5_n_host, &fs.a5_n_host_cap, (size_t)B)) != WDR_OK) return rc;
        if ((rc = grow_dev(&fs.a5_thold_dev, &fs.a5_thold_cap, n_tok)) != WDR_OK) return rc;
        A5Token* th = reinterpret_cast<A5Token*>(fs.a5_tok_host);
        for (int b = 0; b < B; b++) {
            fs.a5_n_host[b] = a5_n[b];
            if (a5_n[b] <= 0) continue;
            const ResultSegment& rs = st->results[a5_seg[b]];
            for (int j = 0; j < a5_n[b]; j++) {
                th[(size_t)b * kDecMaxTokens + j] = A5Token{rs.tokens[j].t0, rs.tokens[j].t1};
                fs.a5_text_host[(size_t)b * kDecMaxTokens + j] = rs.tokens[j].id < v.eot ? 1 : 0;
            }
        }
        // side stream: these kernels touch a few SMs' worth of threads and no tensor-core resources, so they run under the DTW pass
        cudaStream_t cs = st->copy_stream;
        WDR_CUDA_TRY(cudaStreamWaitEvent(cs, fs.ev_energy, 0));
        WDR_CUDA_TRY(cudaMemcpyAsync(fs.a5_tok_dev, fs.a5_tok_host, sizeof(int64_t) * n_tok * 2, cudaMemcpyHostToDevice, cs));
        WDR_CUDA_TRY(cudaMemcpyAsync(fs.a5_text_dev, fs.a5_text_host, n_tok, cudaMemcpyHostToDevice, cs));
        WDR_CUDA_TRY(cudaMemcpyAsync(fs.a5_n_dev, fs.a5_n_host, sizeof(int32_t) * B, cudaMemcpyHostToDevice, cs));
        {
            ProfScope ps(&st->prof, KC_OTHER, cs);
            a5_thold_kernel<<<dim3(8, B), 256, 0, cs>>>(reinterpret_cast<const A5Token*>(fs.a5_tok_dev), fs.a5_text_dev, fs.a5_n_dev, fs.nvalid_dev, fs.energy_dev,
                                                       WDR_CHUNK_SAMPLES, fs.a5_thold_dev);
            WDR_LAUNCH_CHECK();
        }
        {
            ProfScope ps(&st->prof, KC_OTHER, cs);
            a5_adjust_kernel<<<B, kA5Threads, 0, cs>>>(reinterpret_cast<A5Token*>(fs.a5_tok_dev), fs.a5_text_dev, fs.a5_n_dev, fs.nvalid_dev, fs.energy_dev, WDR_CHUNK_SAMPLES,
                                              fs.a5_thold_dev);
            WDR_LAUNCH_CHECK();
        }
        WDR_CUDA_TRY(cudaMemcpyAsync(fs.a5_tok_host, fs.a5_tok_dev, sizeof(int64_t) * n_tok * 2, cudaMemcpyDeviceToHost, cs));
        WDR_CUDA_TRY(cudaEventRecord(fs.ev_energy_done, cs));
    }
    auto apply_a5 = [&]() {  // after the stream has drained: the adjusted times go into the result tokens
        if (!a5_any) return;
        const A5Token* th = reinterpret_cast<const A5Token*>(fs.a5_tok_host);
        for (int b = 0; b < B; b++) {
            if (a5_n[b] <= 0) continue;
            ResultSegment& rs = st->results[a5_seg[b]];
            for (int j = 0; j < a5_n[b]; j++) {
                rs.tokens[j].t0 = th[(size_t)b * kDecMaxTokens + j].t0;
                rs.tokens[j].t1 = th[(size_t)b * kDecMaxTokens + j].t1;
            }
        }
    };
    // ---- DTW token timestamps: teacher-forced pass capturing the alignment heads, then cost + wavefront + backtrace ----
    if (!pend.empty()) {
        const int Ha = (int)ctx->aheads.size();
        std::vector<int64_t> aw_off(B, 0), x_off(B, 0);
        std::vector<int32_t> aw_T(B, 0), aw_A(B, 0);
        std::fill(seq.begin(), seq.end(), v.eot);
        size_t aw_total = 0, x_total = 0, stat_max = 0;
        int max_T = 0, max_path = 1;
        for (auto& pd : pend) {
            const int T_b = (int)pd.dtw_seq.size(), A_b = pd.n_frames / 2;
            aw_T[pd.b] = T_b; aw_A[pd.b] = A_b;
            aw_off[pd.b] = (int64_t)aw_total;
            aw_total += (size_t)Ha * T_b * A_b;
            x_off[pd.b] = (int64_t)x_total;
            x_total += (size_t)(T_b - pd.sot_len - 1) * A_b;
            stat_max = std::max(stat_max, (size_t)Ha * A_b);
            max_T = std::max(max_T, T_b);
            max_path = std::max(max_path, T_b + A_b);
            for (int i = 0; i < T_b; i++) seq[(size_t)pd.b * kDecSeqCap + i] = pd.dtw_seq[i];
        }
        if (aw_total > ws.aw_cap) {
            if (ws.aw) cudaFree(ws.aw);
            ws.aw = nullptr; ws.aw_cap = 0;
            WDR_CUDA_TRY(cudaMalloc(&ws.aw, sizeof(float) * aw_total));
            ws.aw_cap = aw_total;
        }
        if ((rc = grow_dev(&fs.dtw_x, &fs.dtw_x_cap, std::max<size_t>(x_total, 1))) != WDR_OK) return rc;
        if ((rc = grow_dev(&fs.dtw_stat, &fs.dtw_stat_cap, 2 * std::max<size_t>(stat_max, 1))) != WDR_OK) return rc;
        if ((rc = grow_dev(&fs.dtw_path, &fs.dtw_path_cap, (size_t)(2 * max_path + 1) * B)) != WDR_OK) return rc;
        if ((rc = grow_dev(&fs.dtw_wins, &fs.dtw_wins_cap, (size_t)B * sizeof(DtwWindow))) != WDR_OK) return rc;
        if ((rc = grow_pinned(&fs.dtw_path_host, &fs.dtw_path_host_cap, (size_t)(2 * max_path + 1) * B)) != WDR_OK) return rc;
        WDR_CUDA_TRY(cudaMemcpyAsync(ws.seq, seq.data(), sizeof(int32_t) * seq.size(), cudaMemcpyHostToDevice, s));
        WDR_CUDA_TRY(cudaMemcpyAsync(ws.aw_off, aw_off.data(), sizeof(int64_t) * B, cudaMemcpyHostToDevice, s));
        WDR_CUDA_TRY(cudaMemcpyAsync(ws.aw_T, aw_T.data(), sizeof(int32_t) * B, cudaMemcpyHostToDevice, s));
        WDR_CUDA_TRY(cudaMemcpyAsync(ws.aw_A, aw_A.data(), sizeof(int32_t) * B, cudaMemcpyHostToDevice, s));
        static const bool stepwise = getenv("WDR_DTW_STEPWISE") != nullptr;  // bring-up aid: the one-token-per-launch pass
        if (stepwise) {
            for (int i = 0; i < max_T; i++)
                if ((rc = decoder_step(ctx, ws, B, i, false, DEC_MODE_DTW, s, &st->prof)) != WDR_OK) return rc;
        } else {
            static const bool prof_range = getenv("WDR_PROFILE_DTW_PASS") != nullptr;  // ncu --profile-from-start off: capture this pass only
            if (prof_range) { cudaStreamSynchronize(s); cudaProfilerStart(); }
            rc = decoder_dtw_pass(ctx, ws, st->dtwp, B, aw_T.data(), s, &st->prof);
            if (prof_range) { cudaStreamSynchronize(s); cudaProfilerStop(); }
            if (rc != WDR_OK) return rc;
        }
        WDR_CUDA_TRY(cudaEventRecord(fs.ev_phase[4], s));
        std::vector<DtwWindow> wins;
        std::vector<int> win_b;
        for (auto& pd : pend) {
            const int T_b = aw_T[pd.b], A_b = aw_A[pd.b], N = T_b - pd.sot_len - 1;
            if (N <= 0 || A_b <= 3) continue;
            {
                ProfScope ps(&st->prof, KC_DTW, s);
                if ((rc = dtw_cost_dev(ws.aw + aw_off[pd.b], Ha, T_b, A_b, pd.sot_len, 7, fs.dtw_stat, fs.dtw_stat + stat_max, fs.dtw_x + x_off[pd.b], s)) != WDR_OK) return rc;
            }
            DtwWindow dw;
            dw.x_off = x_off[pd.b]; dw.N = N; dw.M = A_b; dw.tr_off = 0;
            wins.push_back(dw);
            win_b.push_back((int)(&pd - &pend[0]));
        }
        if (!wins.empty()) {
            int32_t* ti = fs.dtw_path;
            int32_t* tj = ti + (size_t)max_path * B;
            int32_t* pl = tj + (size_t)max_path * B;
            {
                ProfScope ps(&st->prof, KC_DTW, s);
                if ((rc = dtw_run(fs.dtw_x, wins, ti, tj, pl, max_path, nullptr, nullptr, s, fs.dtw_wins)) != WDR_OK) return rc;
            }
            const size_t nw = wins.size();
            int32_t* h_ti = fs.dtw_path_host;  // pinned: the three reads below are true async copies
            int32_t* h_tj = h_ti + nw * max_path;
            int32_t* h_pl = h_tj + nw * max_path;
            WDR_CUDA_TRY(cudaMemcpyAsync(h_ti, ti, sizeof(int32_t) * nw * max_path, cudaMemcpyDeviceToHost, s));
            WDR_CUDA_TRY(cudaMemcpyAsync(h_tj, tj, sizeof(int32_t) * nw * max_path, cudaMemcpyDeviceToHost, s));
            WDR_CUDA_TRY(cudaMemcpyAsync(h_pl, pl, sizeof(int32_t) * nw, cudaMemcpyDeviceToHost, s));
            run_post();  // host work under the queued GPU work
            WDR_CUDA_TRY(cudaStreamSynchronize(s));
            for (size_t k = 0; k < nw; k++) {
                const Pending& pd = pend[win_b[k]];
                ResultSegment& seg = st->results[pd.seg];
                const int seek = sw ? sw->seek : 0;
                const int32_t* a_ti = h_ti + k * max_path;
                const int32_t* a_tj = h_tj + k * max_path;
                if (h_pl[k] < 0) { set_error("dtw backtrace did not terminate"); return WDR_ERR_CUDA; }
                int last_v = 0;
                size_t tix = 0;
                for (int i = 0; i < h_pl[k]; i++) {
                    const int vv = a_ti[i];
                    if (vv != last_v) {
                        const int64_t ts = (int64_t)a_tj[i] * 2 + seek;
                        last_v = vv;
                        while (tix < seg.tokens.size() && !(seg.tokens[tix].id < v.eot)) tix++;
                        if (tix >= seg.tokens.size()) break;
                        seg.tokens[tix].t_dtw = ts;
                        tix++;
                    }
                }
            }
        } else {
            run_post();
            WDR_CUDA_TRY(cudaStreamSynchronize(s));
        }
    }
    run_post();  // no DTW: nothing was queued above
    if (a5_any) WDR_CUDA_TRY(cudaEventSynchronize(fs.ev_energy_done));
    apply_a5();
    {   // cross-attention launch / live-window counters of this group (stepwise DTW passes included)
        if (!fs.cross_stats_host) WDR_CUDA_TRY(cudaMallocHost(reinterpret_cast<void**>(&fs.cross_stats_host), sizeof(unsigned long long) * 2));
        WDR_CUDA_TRY(cudaMemcpyAsync(fs.cross_stats_host, ws.cross_stats, sizeof(unsigned long long) * 2, cudaMemcpyDeviceToHost, s));
        WDR_CUDA_TRY(cudaStreamSynchronize(s));
        fs.cross_launches += (int64_t)fs.cross_stats_host[0];
        fs.cross_live += (int64_t)fs.cross_stats_host[1];
    }
    {   // phase times of this group (the stream is idle here: every branch above ended with a synchronize)
        const bool dtw = !pend.empty();
        if (!dtw) WDR_CUDA_TRY(cudaEventRecord(fs.ev_phase[4], s));
        WDR_CUDA_TRY(cudaEventRecord(fs.ev_phase[5], s));
        WDR_CUDA_TRY(cudaEventSynchronize(fs.ev_phase[5]));
        for (int i = 0; i < 5; i++) {
            float t = 0.0f;
            if (cudaEventElapsedTime(&t, fs.ev_phase[i], fs.ev_phase[i + 1]) == cudaSuccess) fs.phase_ms[i] += t;
        }
    }
    return WDR_OK;
}

// whisper_full_with_state on a buffer longer than 30 s: the seek loop of SURVEY A.4 (what the crate runs when VAD and
// diarization are off: one SpeechSegment holding the whole file, reference src/engine.rs:124-134, src/transcribe.rs:389).
// The mel of the whole buffer is computed once and normalised with its global maximum; windows are decoded one at a time
// (window k+1 starts where window k's last timestamp token says, and is conditioned on window k's tokens), so this mode
// is sequential by construction: replicas only, no sharding.
template <typename In>
static int full_sequential(wdr_context* ctx, wdr_state* st, const wdr_full_params& p, int lang_id, const In* pcm_host, int n) {
    int rc = ensure_device(ctx->device);
    if (rc != WDR_OK) return rc;
    st->results.clear();
    st->chunk_info.clear();
    st->lang_id = lang_id;
    FullScratch& fs = st->full;
    for (auto& v : fs.phase_ms) v = 0.0;
    fs.decode_steps = 0;
    fs.cross_launches = fs.cross_live = 0;
    if (!fs.ev_energy) {
        WDR_CUDA_TRY(cudaEventCreateWithFlags(&fs.ev_energy, cudaEventDisableTiming));
        WDR_CUDA_TRY(cudaEventCreateWithFlags(&fs.ev_energy_done, cudaEventDisableTiming));
        WDR_CUDA_TRY(cudaEventCreateWithFlags(&fs.ev_h2d, cudaEventDisableTiming));
        WDR_CUDA_TRY(cudaMallocHost(reinterpret_cast<void**>(&fs.done_host), sizeof(int32_t)));
    }
    cudaStream_t s = st->stream;
    const int n_len = (n + WDR_CHUNK_SAMPLES) / WDR_HOP;
    const int n_mel = ctx->arch.n_mel;
    DevBuf<In> d_pcm;
    DevBuf<float> d_mel, d_max, d_energy;
    WDR_CUDA_TRY(d_pcm.alloc((size_t)n));
    WDR_CUDA_TRY(d_mel.alloc((size_t)n_mel * n_len));
    WDR_CUDA_TRY(d_max.alloc(1));
    WDR_CUDA_TRY(cudaMemcpyAsync(d_pcm.p, pcm_host, sizeof(In) * (size_t)n, cudaMemcpyHostToDevice, s));
    if ((rc = mel_launch<In>(ctx->mel, d_pcm.p, 0, nullptr, n, 1, n_len, 0, d_mel.p, d_max.p, s)) != WDR_OK) return rc;
    std::vector<float> energy;
    if (p.token_timestamps) {
        WDR_CUDA_TRY(d_energy.alloc((size_t)n));
        energy.resize((size_t)n);
        energy_kernel<In><<<dim3(148 * 8, 1), 256, 0, s>>>(d_pcm.p, 0, nullptr, n, 32, d_energy.p, 0);
        WDR_LAUNCH_CHECK();
        WDR_CUDA_TRY(cudaMemcpyAsync(energy.data(), d_energy.p, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, s));
    }
    WDR_CUDA_TRY(cudaStreamSynchronize(s));
    const int seek_end = 1 + (n + 200 - WDR_N_FFT) / WDR_HOP;  // mel.n_len_org
    int seek = 0;
    int64_t st3[3] = {0, 0, 0};
    std::vector<int32_t> prompt_past;  // no_context = true (the crate never clears it): starts empty for every call
    if (p.prompt_tokens && p.prompt_n_tokens > 0) prompt_past.assign(p.prompt_tokens, p.prompt_tokens + p.prompt_n_tokens);
    const Vocab v = ctx_vocab(ctx);
    if (seek_end < seek + kDeltaMin) return WDR_OK;
    while (seek + kDeltaMin < seek_end) {
        SeqWindow sw;
        sw.mel_dev = d_mel.p; sw.n_len = n_len; sw.max_dev = d_max.p; sw.seek = seek; sw.seek_end = seek_end;
        sw.energy_host = energy.empty() ? nullptr : energy.data(); sw.n_samples = n; sw.st3 = st3; sw.prompt_past = &prompt_past;
        std::vector<int32_t> result_ids;
        sw.result_ids = &result_ids;
        rc = full_group<In>(ctx, st, p, lang_id, nullptr, 0, nullptr, 0, 1, &sw);
        if (rc != WDR_OK) return rc;
        const ChunkInfo& ci = st->chunk_info.back();
        if (lang_id < 0) lang_id = ci.lang_id;  // whisper_full detects once, on the first window
        if (p.detect_language) return WDR_OK;
        // whisper_full: prompt_past.clear(); if the winning prompt began with [PREV] (it does only below temperature 0.5) the part
        // of the old past that went into it is kept; then tokens_cur[0 .. result_len) are appended unless the window is a
        // no-speech window — whether or not a segment came out of them
        {
            std::vector<int32_t> next;
            if (!prompt_past.empty() && p.n_max_text_ctx > 0 && ci.temperature < 0.5f) {
                const int n_take = std::min(std::min(p.n_max_text_ctx, WDR_TEXT_CTX / 2), (int)prompt_past.size());
                next.assign(prompt_past.end() - n_take, prompt_past.end());
            }
            next.insert(next.end(), result_ids.begin(), result_ids.end());
            prompt_past.swap(next);
        }
        if (p.progress_callback) p.progress_callback(ctx, st, (int)std::min<int64_t>(100, 100ll * (seek + ci.seek_delta) / std::max(1, seek_end)), p.progress_callback_user_data);
        if (ci.seek_delta <= 0) break;  // no progress: stop rather than spin (whisper.cpp cannot get here with a positive delta_min)
        seek += ci.seek_delta;
    }
    (void)v;
    return WDR_OK;
}

// Shared progress accounting of one full call (lanes report under the mutex; the callback never runs concurrently).
struct FullProgress {
    std::mutex mu;
    int done = 0, total = 0;
    wdr_context* ctx = nullptr;
    wdr_state* owner = nullptr;
    void add(const wdr_full_params& p, int n) {
        std::lock_guard<std::mutex> lk(mu);
        done += n;
        if (p.progress_callback && total > 0) p.progress_callback(ctx, owner, (int)(100ll * done / total), p.progress_callback_user_data);
    }
};

// Chunks [c_begin, c_end) of the call on one lane (its own streams and workspaces), in groups of <= 128 windows.
template <typename In>
static int full_range(wdr_context* ctx, wdr_state* st, const wdr_full_params& p, int lang_id, const In* pcm, int64_t chunk_stride,
                      const int32_t* n_valid, int c_begin, int c_end, bool pcm_on_device, FullProgress* prog) {
    int rc = ensure_device(ctx->device);  // the current device is per host thread
    if (rc != WDR_OK) return rc;
    st->results.clear();
    st->chunk_info.clear();
    st->lang_id = lang_id;
    FullScratch& fs = st->full;
    for (auto& v : fs.phase_ms) v = 0.0;
    fs.decode_steps = 0;
    fs.cross_launches = fs.cross_live = 0;
    if (!fs.ev_energy) {
        WDR_CUDA_TRY(cudaEventCreateWithFlags(&fs.ev_energy, cudaEventDisableTiming));
        WDR_CUDA_TRY(cudaEventCreateWithFlags(&fs.ev_energy_done, cudaEventDisableTiming));
        WDR_CUDA_TRY(cudaEventCreateWithFlags(&fs.ev_h2d, cudaEventDisableTiming));
        WDR_CUDA_TRY(cudaMallocHost(reinterpret_cast<void**>(&fs.done_host), sizeof(int32_t)));
    }
    // windows per decode group: bounded by the cross-KV memory (128 windows) and by the row capacity (windows x beams <= 640)
    const int group_max = std::min(kDecMaxWindows, kDecMaxRows / ((p.strategy == WDR_SAMPLING_BEAM_SEARCH && p.beam_size > 1) ? p.beam_size : 1));
    for (int c0 = c_begin; c0 < c_end; c0 += group_max) {
        const int B = std::min(group_max, c_end - c0);
        if (pcm_on_device) {
            rc = full_group<In>(ctx, st, p, lang_id, pcm + (size_t)c0 * chunk_stride, chunk_stride, n_valid, c0, B);
            if (rc != WDR_OK) return rc;
            prog->add(p, B);
            continue;
        }
        // stage this group's PCM: rows of up to 480000 samples at a fixed device stride
        const size_t need = (size_t)B * WDR_CHUNK_SAMPLES * sizeof(In);
        if (need > fs.pcm_cap) {
            if (fs.pcm_dev) cudaFree(fs.pcm_dev);
            fs.pcm_dev = nullptr; fs.pcm_cap = 0;
            WDR_CUDA_TRY(cudaMalloc(&fs.pcm_dev, need));
            fs.pcm_cap = need;
        }
        In* pcm_dev = reinterpret_cast<In*>(fs.pcm_dev);
        for (int b = 0; b < B; b++) {
            const int n = n_valid ? n_valid[c0 + b] : WDR_CHUNK_SAMPLES;
            if (n > 0)
                WDR_CUDA_TRY(cudaMemcpyAsync(pcm_dev + (size_t)b * WDR_CHUNK_SAMPLES, pcm + (size_t)(c0 + b) * chunk_stride, sizeof(In) * (size_t)n,
                                             cudaMemcpyHostToDevice, st->stream));
        }
        rc = full_group<In>(ctx, st, p, lang_id, pcm_dev, WDR_CHUNK_SAMPLES, n_valid, c0, B);
        if (rc != WDR_OK) return rc;
        prog->add(p, B);
    }
    return WDR_OK;
}

// Lanes: the windows of one call can be cut into contiguous ranges, each driven by its own host thread on its own streams and
// workspaces (one lane's tensor-bound encoder / latency-bound decode chain under another's HBM-bound cross-attention).
// Windows are independent (sharded mode) and every kernel's result for a window is independent of which other
// windows share its launch, so the output is identical for any lane count.  Measured on B200 (large-v3, 120 windows): 2 lanes
// are no faster than 1 — each lane's weight-streaming GEMMs still occupy every SM's shared memory and a half-full 128-row tile
// costs what a full one does — so the default is ONE lane; the knob stays for many-small-window calls.  A low-footprint
// cross-attention for multi-lane calls (2 persistent CTAs per SM at 48 registers, so that the other lane's GEMM CTA fits next to
// them) was measured as well: 2 lanes 2568 ms vs 2190 ms for one lane (each lane's decode 2085 vs 1700 ms), so it was dropped.
static int default_lanes() {
    static const int n = [] {
        const char* e = getenv("WDR_LANES");
        const int v = e ? atoi(e) : 1;
        return v < 1 ? 1 : (v > 8 ? 8 : v);
    }();
    return n;
}
constexpr int kMinWindowsPerLane = 8;

template <typename In>
static int full_batch_impl(wdr_context* ctx, wdr_state* st, const wdr_full_params& p, const In* pcm, int64_t chunk_stride,
                           const int32_t* n_valid, int n_chunks, bool pcm_on_device = false) {
    clear_error();
    WDR_REQUIRE(ctx && st && st->ctx == ctx && n_chunks >= 0, "bad arguments");
    WDR_REQUIRE(n_chunks == 0 || (pcm && chunk_stride >= 0), "bad arguments");
    int rc = ensure_device(ctx->device);
    if (rc != WDR_OK) return rc;
    int lang_id = 0;
    if ((rc = validate_params(ctx, p, &lang_id)) != WDR_OK) return rc;
    for (int b = 0; b < n_chunks; b++) {
        const int n = n_valid ? n_valid[b] : WDR_CHUNK_SAMPLES;
        WDR_REQUIRE(n >= 0 && n <= WDR_CHUNK_SAMPLES, "each buffer must hold 0..480000 samples (longer audio: split into 30 s chunks)");
    }
    if (pcm_on_device) WDR_REQUIRE(chunk_stride >= WDR_CHUNK_SAMPLES, "device PCM needs chunk_stride >= 480000");
    FullProgress prog;
    prog.total = n_chunks; prog.ctx = ctx; prog.owner = st;
    const int want = st->n_lanes > 0 ? st->n_lanes : default_lanes();
    const int G = std::max(1, std::min(want, n_chunks / kMinWindowsPerLane));
    if (G == 1) return full_range<In>(ctx, st, p, lang_id, pcm, chunk_stride, n_valid, 0, n_chunks, pcm_on_device, &prog);
    while ((int)st->lanes.size() < G - 1) {
        wdr_state* ln = wdr_init_state(ctx);
        if (!ln) return WDR_ERR_CUDA;
        st->lanes.push_back(ln);
    }
    std::vector<int> bounds(G + 1);
    for (int g = 0; g <= G; g++) bounds[g] = (int)((int64_t)n_chunks * g / G);
    std::vector<int> rcs(G, WDR_OK);
    std::vector<std::string> errs(G);
    std::vector<std::thread> th;
    auto run = [&](int g) {
        wdr_state* ln = g == 0 ? st : st->lanes[g - 1];
        ln->prof.enabled = st->prof.enabled;
        clear_error();
        rcs[g] = full_range<In>(ctx, ln, p, lang_id, pcm, chunk_stride, n_valid, bounds[g], bounds[g + 1], pcm_on_device, &prog);
        if (rcs[g] != WDR_OK) errs[g] = wdr_last_error();
    };
    for (int g = 1; g < G; g++) th.emplace_back(run, g);
    run(0);
    for (auto& t : th) t.join();
    for (int g = 0; g < G; g++)
        if (rcs[g] != WDR_OK) {
            st->results.clear();
            st->chunk_info.clear();
            set_error("%s", errs[g].c_str());
            return rcs[g];
        }
    for (int g = 1; g < G; g++) {  // lane ranges are contiguous and ascending: appending keeps chunk order
        wdr_state* ln = st->lanes[g - 1];
        for (auto& r : ln->results) st->results.push_back(std::move(r));
        st->chunk_info.insert(st->chunk_info.end(), ln->chunk_info.begin(), ln->chunk_info.end());
        ln->results.clear();
        ln->chunk_info.clear();
    }
    return WDR_OK;
}

template <typename In>
static int full_long(wdr_context* ctx, wdr_state* st, const wdr_full_params& p, const In* pcm, int n) {
    WDR_REQUIRE(ctx && st && st->ctx == ctx && pcm, "bad arguments");
    int lang_id = 0;
    int rc = validate_params(ctx, p, &lang_id);
    if (rc != WDR_OK) return rc;
    return full_sequential<In>(ctx, st, p, lang_id, pcm, n);
}

}  // namespace wdr

using namespace wdr;

extern "C" wdr_full_params wdr_full_default_params(int strategy) {
    wdr_full_params p;
    memset(&p, 0, sizeof(p));
    p.strategy = strategy;
    p.n_threads = 4;
    p.n_max_text_ctx = 16384;
    p.no_context = 1;
    p.print_progress = 1;
    p.print_timestamps = 1;
    p.thold_pt = 0.01f;
    p.thold_ptsum = 0.01f;
    p.language = "en";
    p.suppress_blank = 1;
    p.temperature = 0.0f;
    p.max_initial_ts = 1.0f;
    p.length_penalty = -1.0f;
    p.temperature_inc = 0.2f;  // == whisper_full_default_params: a host that takes the defaults gets the fallback ladder
    p.entropy_thold = 2.4f;
    p.logprob_thold = -1.0f;
    p.no_speech_thold = 0.6f;
    p.greedy_best_of = strategy == WDR_SAMPLING_GREEDY ? 5 : -1;
    p.beam_size = strategy == WDR_SAMPLING_BEAM_SEARCH ? 5 : -1;
    p.beam_patience = -1.0f;
    return p;
}

extern "C" int wdr_full_with_state(wdr_context* ctx, wdr_state* st, wdr_full_params p, const float* pcm, int n) {
    clear_error();
    std::vector<int32_t> prompt_store;
    resolve_initial_prompt(ctx, p, prompt_store);
    WDR_REQUIRE(n >= 0, "negative sample count");
    if (n > WDR_CHUNK_SAMPLES) return full_long<float>(ctx, st, p, pcm, n);
    const int32_t nv = n;
    return full_batch_impl<float>(ctx, st, p, pcm, WDR_CHUNK_SAMPLES, &nv, 1);
}
extern "C" int wdr_full_with_state_i16(wdr_context* ctx, wdr_state* st, wdr_full_params p, const int16_t* pcm, int n) {
    clear_error();
    std::vector<int32_t> prompt_store;
    resolve_initial_prompt(ctx, p, prompt_store);
    WDR_REQUIRE(n >= 0, "negative sample count");
    if (n > WDR_CHUNK_SAMPLES) return full_long<int16_t>(ctx, st, p, pcm, n);
    const int32_t nv = n;
    return full_batch_impl<int16_t>(ctx, st, p, pcm, WDR_CHUNK_SAMPLES, &nv, 1);
}
extern "C" int wdr_full_batch_i16(wdr_context* ctx, wdr_state* st, wdr_full_params p, const int16_t* pcm, int64_t chunk_stride,
                                  const int32_t* n_valid, int n_chunks) {
    std::vector<int32_t> prompt_store;
    resolve_initial_prompt(ctx, p, prompt_store);
    return full_batch_impl<int16_t>(ctx, st, p, pcm, chunk_stride, n_valid, n_chunks);
}

extern "C" int wdr_full_batch_i16_dev(wdr_context* ctx, wdr_state* st, wdr_full_params p, const int16_t* pcm_dev, int64_t chunk_stride,
                                      const int32_t* n_valid, int n_chunks) {
    std::vector<int32_t> prompt_store;
    resolve_initial_prompt(ctx, p, prompt_store);
    return full_batch_impl<int16_t>(ctx, st, p, pcm_dev, chunk_stride, n_valid, n_chunks, true);
}

#define SEG_OR(ret)                                                                  \
    if (!st || i < 0 || i >= (int)st->results.size()) { set_error("segment index out of range"); return ret; } \
    const ResultSegment& seg = st->results[i];

extern "C" int wdr_full_n_segments_from_state(wdr_state* st) { return st ? (int)st->results.size() : 0; }
extern "C" int wdr_full_get_segment_chunk_from_state(wdr_state* st, int i) { SEG_OR(-1) return seg.chunk; }
extern "C" int64_t wdr_full_get_segment_t0_from_state(wdr_state* st, int i) { SEG_OR(-1) return seg.t0; }
extern "C" int64_t wdr_full_get_segment_t1_from_state(wdr_state* st, int i) { SEG_OR(-1) return seg.t1; }
extern "C" const char* wdr_full_get_segment_text_from_state(wdr_state* st, int i) { SEG_OR(nullptr) return seg.text.c_str(); }
extern "C" float wdr_full_get_segment_no_speech_prob_from_state(wdr_state* st, int i) { SEG_OR(0.0f) return seg.no_speech_prob; }
extern "C" int wdr_full_n_tokens_from_state(wdr_state* st, int i) { SEG_OR(-1) return (int)seg.tokens.size(); }
extern "C" int32_t wdr_full_get_token_id_from_state(wdr_state* st, int i, int j) {
    SEG_OR(-1)
    if (j < 0 || j >= (int)seg.tokens.size()) { set_error("token index out of range"); return -1; }
    return seg.tokens[j].id;
}
extern "C" const char* wdr_full_get_token_text_from_state(wdr_context*, wdr_state* st, int i, int j) {
    SEG_OR(nullptr)
    if (j < 0 || j >= (int)seg.token_text.size()) { set_error("token index out of range"); return nullptr; }
    return seg.token_text[j].c_str();
}
extern "C" wdr_token_data wdr_full_get_token_data_from_state(wdr_state* st, int i, int j) {
    wdr_token_data z;
    memset(&z, 0, sizeof(z));
    z.t0 = z.t1 = z.t_dtw = -1;
    SEG_OR(z)
    if (j < 0 || j >= (int)seg.tokens.size()) { set_error("token index out of range"); return z; }
    return seg.tokens[j];
}
extern "C" int wdr_full_get_phase_ms(wdr_state* st, double* ms, int32_t* decode_steps) {
    clear_error();
    WDR_REQUIRE(st && ms, "bad arguments");
    for (int i = 0; i < 5; i++) ms[i] = st->full.phase_ms[i];
    int steps = st->full.decode_steps;
    for (auto ln : st->lanes) {
        for (int i = 0; i < 5; i++) ms[i] += ln->full.phase_ms[i];
        steps += ln->full.decode_steps;
    }
    if (decode_steps) *decode_steps = steps;
    return WDR_OK;
}
extern "C" int wdr_full_get_cross_attn_stats(wdr_state* st, int64_t* launches, int64_t* live_windows) {
    clear_error();
    WDR_REQUIRE(st && launches && live_windows, "bad arguments");
    *launches = st->full.cross_launches;
    *live_windows = st->full.cross_live;
    for (auto ln : st->lanes) { *launches += ln->full.cross_launches; *live_windows += ln->full.cross_live; }
    return WDR_OK;
}
extern "C" int wdr_full_lang_id_from_state(wdr_state* st) { return st ? st->lang_id : -1; }
extern "C" int wdr_full_get_chunk_lang_id_from_state(wdr_state* st, int i) {
    if (!st || i < 0 || i >= (int)st->chunk_info.size()) { set_error("chunk index out of range"); return -1; }
    return st->chunk_info[i].lang_id;
}
extern "C" const char* wdr_lang_str(int id) { return (id >= 0 && id < 100) ? kLangs[id] : nullptr; }
extern "C" int wdr_lang_id(const char* lang) { return lang_id_from_str(lang); }
extern "C" const char* wdr_token_to_str(wdr_context* ctx, int32_t token) {
    static thread_local std::string buf;
    if (!ctx || token < 0 || token >= ctx->arch.n_vocab) return nullptr;
    buf = token_text(ctx_vocab(ctx), token);
    return buf.c_str();
}
extern "C" int wdr_full_get_chunk_info_from_state(wdr_state* st, int i, int32_t* info, float* nsp) {
    clear_error();
    WDR_REQUIRE(st && info && i >= 0 && i < (int)st->chunk_info.size(), "chunk index out of range");
    const ChunkInfo& c = st->chunk_info[i];
    info[0] = c.seek_delta; info[1] = c.failed; info[2] = c.completed; info[3] = c.n_sampled; info[4] = c.has_ts; info[5] = c.result_len;
    info[6] = c.seek_end; info[7] = c.n_segments;
    if (nsp) *nsp = c.no_speech_prob;
    return WDR_OK;
}

extern "C" int wdr_sample_discrete(const float* logprobs, int n, uint32_t seed, int n_draws, int32_t* ids) {
    clear_error();
    WDR_REQUIRE(logprobs && n > 0 && n_draws >= 0 && (n_draws == 0 || ids), "bad arguments");
    std::mt19937 rng(seed);
    std::vector<float> probs;
    for (int i = 0; i < n_draws; i++) ids[i] = sample_discrete(logprobs, n, rng, probs);
    return WDR_OK;
}

extern "C" float wdr_full_get_chunk_temperature_from_state(wdr_state* st, int i) {
    clear_error();
    if (!st || i < 0 || i >= (int)st->chunk_info.size()) { set_error("chunk index out of range"); return -1.0f; }
    return st->chunk_info[i].temperature;
}

extern "C" int wdr_decode_teacher_forced(wdr_context* ctx, wdr_state* st, const float* enc, int n_chunks, const int32_t* seq_in, int n_seq,
                                         float* logits_out, float* aheads_out) {
    clear_error();
    WDR_REQUIRE(ctx && st && st->ctx == ctx && seq_in && n_chunks > 0 && n_chunks <= kDecMaxWindows && n_seq > 0 && n_seq <= kDecSeqCap, "bad arguments");
    WDR_REQUIRE(!aheads_out || !ctx->aheads.empty(), "alignment heads need a context created with dtw_token_timestamps");
    int rc = ensure_device(ctx->device);
    if (rc != WDR_OK) return rc;
    DecoderWorkspace& ws = st->dec;
    const int B = n_chunks, d = ctx->arch.d, nv = ctx->arch.n_vocab;
    cudaStream_t s = st->stream;
    if (enc) {
        if ((rc = ws.reserve(ctx, B)) != WDR_OK) return rc;
        const size_t n = (size_t)B * WDR_AUDIO_CTX * d;
        DevBuf<float> tmp;
        WDR_CUDA_TRY(tmp.alloc(n));
        WDR_CUDA_TRY(cudaMemcpyAsync(tmp.p, enc, sizeof(float) * n, cudaMemcpyHostToDevice, s));
        f32_to_bf16_kernel<<<148 * 4, 256, 0, s>>>(tmp.p, ws.enc_bf16, (int64_t)n);
        WDR_LAUNCH_CHECK();
        if ((rc = decoder_cross_kv(ctx, ws, B, s, &st->prof)) != WDR_OK) return rc;
        WDR_CUDA_TRY(cudaStreamSynchronize(s));
    }
    WDR_REQUIRE(ws.cap_W >= B, "no encoder output in the state for that many windows");
    const Vocab v = make_vocab(nv);
    std::vector<int32_t> seq((size_t)B * kDecSeqCap, v.eot);
    for (int b = 0; b < B; b++)
        for (int i = 0; i < n_seq; i++) seq[(size_t)b * kDecSeqCap + i] = seq_in[(size_t)b * n_seq + i];
    WDR_CUDA_TRY(cudaMemcpyAsync(ws.seq, seq.data(), sizeof(int32_t) * seq.size(), cudaMemcpyHostToDevice, s));
    const int Ha = (int)ctx->aheads.size();
    if (aheads_out) {
        const size_t per = (size_t)Ha * n_seq * WDR_AUDIO_CTX;
        if (per * B > ws.aw_cap) {
            if (ws.aw) cudaFree(ws.aw);
            ws.aw = nullptr; ws.aw_cap = 0;
            WDR_CUDA_TRY(cudaMalloc(&ws.aw, sizeof(float) * per * B));
            ws.aw_cap = per * B;
        }
        std::vector<int64_t> off(B);
        std::vector<int32_t> T(B, n_seq), A(B, WDR_AUDIO_CTX);
        for (int b = 0; b < B; b++) off[b] = (int64_t)(per * b);
        WDR_CUDA_TRY(cudaMemcpyAsync(ws.aw_off, off.data(), sizeof(int64_t) * B, cudaMemcpyHostToDevice, s));
        WDR_CUDA_TRY(cudaMemcpyAsync(ws.aw_T, T.data(), sizeof(int32_t) * B, cudaMemcpyHostToDevice, s));
        WDR_CUDA_TRY(cudaMemcpyAsync(ws.aw_A, A.data(), sizeof(int32_t) * B, cudaMemcpyHostToDevice, s));
    }
    for (int i = 0; i < n_seq; i++) {
        if ((rc = decoder_step(ctx, ws, B, i, logits_out != nullptr, aheads_out ? DEC_MODE_DTW : DEC_MODE_FORCED, s, &st->prof)) != WDR_OK) return rc;
        if (logits_out)
            WDR_CUDA_TRY(cudaMemcpy2DAsync(logits_out + (size_t)i * nv, sizeof(float) * (size_t)n_seq * nv, ws.logits, sizeof(float) * ws.ldv,
                                           sizeof(float) * nv, B, cudaMemcpyDeviceToHost, s));
    }
    if (aheads_out)
        WDR_CUDA_TRY(cudaMemcpyAsync(aheads_out, ws.aw, sizeof(float) * (size_t)B * Ha * n_seq * WDR_AUDIO_CTX, cudaMemcpyDeviceToHost, s));
    WDR_CUDA_TRY(cudaStreamSynchronize(s));
    return WDR_OK;
}

// Bring-up aid (not part of the reference surface): copies a decoder workspace buffer to the host as fp32.
// which: 0 x [B][d] | 1 h | 2 att | 3 ff [B][4d] | 4 first `count` elements of layer-0 cross K|V | 5 split-K partial workspace
extern "C" int wdr_debug_decoder_read(wdr_state* st, int which, float* out, int64_t count) {
    clear_error();
    WDR_REQUIRE(st && out && count > 0, "bad arguments");
    DecoderWorkspace& ws = st->dec;
    WDR_REQUIRE(ws.cap_B > 0, "no decoder workspace yet");
    WDR_CUDA_TRY(cudaStreamSynchronize(st->stream));
    if (which == 0 || which == 5) {
        WDR_CUDA_TRY(cudaMemcpy(out, which == 0 ? ws.x : ws.part, sizeof(float) * count, cudaMemcpyDeviceToHost));
        return WDR_OK;
    }
    const __nv_bfloat16* src = which == 1 ? ws.h : which == 2 ? ws.att : which == 3 ? ws.ff : ws.ckv[0];
    std::vector<__nv_bfloat16> tmp((size_t)count);
    WDR_CUDA_TRY(cudaMemcpy(tmp.data(), src, sizeof(__nv_bfloat16) * count, cudaMemcpyDeviceToHost));
    for (int64_t i = 0; i < count; i++) out[i] = __bfloat162float(tmp[i]);
    return WDR_OK;
}
