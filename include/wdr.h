/*
 * wdr.h — C ABI of libwdr_b200: the B200-native (CUDA sm_100a) replacement for the native compute
 * that tmoroney/whisper-diarize-rs reaches through whisper-rs (whisper.cpp) and pyannote-rs
 * (ONNX Runtime + kaldi-native-fbank) behind Engine::transcribe_audio (reference src/engine.rs:65-200).
 *
 * Every entry point names the reference interface it replaces (reference file:line = the call site in
 * the crate; "whisper.h"/"pyannote-rs" = the un-vendored upstream symbol that call resolves to).
 * Plain pointers and sizes only.  Return convention: 0 (or a non-negative count) = ok, negative = error
 * (see WDR_ERR_*); constructors return NULL on failure and never abort (the crate wraps them in
 * catch_unwind, src/transcribe.rs:153).  There is no CPU fallback: without a CUDA device every compute
 * call returns WDR_ERR_NO_DEVICE.
 *
 * Pointer residency: functions without a suffix take HOST pointers and do their own H2D/D2H on the
 * library's stream; functions ending in _dev take DEVICE pointers plus a cudaStream_t (passed as
 * void*) and never synchronise — those are the ones bench.py times with inputs resident in HBM.
 */
#ifndef WDR_H
#define WDR_H

#include <stddef.h>
#include <stdint.h>
#include <stdbool.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WDR_OK 0
#define WDR_ERR_INVALID (-1)     /* bad argument */
#define WDR_ERR_NO_DEVICE (-2)   /* no CUDA device / driver: there is no CPU path */
#define WDR_ERR_CUDA (-3)        /* a CUDA runtime call or kernel failed; see wdr_last_error() */
#define WDR_ERR_OOM (-4)
#define WDR_ERR_ABORTED (-5)     /* abort callback returned true (src/transcribe.rs:348-350) */
#define WDR_ERR_TOO_SHORT (-6)   /* e.g. fbank yields 0 frames (src/transcribe.rs:466-476 maps it to speaker "?") */
#define WDR_ERR_UNSUPPORTED (-7)

#define WDR_SAMPLE_RATE 16000
#define WDR_N_FFT 400
#define WDR_HOP 160
#define WDR_CHUNK_SAMPLES 480000 /* 30 s */
#define WDR_CHUNK_FRAMES 3000    /* mel frames per 30 s window */
#define WDR_AUDIO_CTX 1500       /* encoder positions per window */
#define WDR_TEXT_CTX 448

/* ---- library-wide ---------------------------------------------------------------------------- */
const char* wdr_version(void);
/* Thread-local description of the last error on this thread ("" if none). */
const char* wdr_last_error(void);
/* Number of CUDA devices visible (0 if none; never negative). */
int wdr_device_count(void);
/* Replaces whisper_log_set (whisper_rs::install_logging_hooks, reference examples/test.rs:6). */
typedef void (*wdr_log_callback)(int level, const char* text, void* user_data);
void wdr_log_set(wdr_log_callback cb, void* user_data);
/* Count of kernels this library has launched in this process (bench.py's gpu_launches). */
uint64_t wdr_launch_count(void);

/* ---- audio (reference src/audio.rs:4-24, src/transcribe.rs:380-381, src/vad.rs:11-12) ---------- */
/* whisper_rs::convert_integer_to_float_audio: out[i] = in[i] / 32768.0f, exact. Device kernel. */
int wdr_convert_integer_to_float_audio(const int16_t* pcm, int n, float* out);
/* Resample interleaved int16 PCM at any common rate to 16 kHz mono (north-star piece 1; the reference's audio::read_wav,
 * src/audio.rs:9-20, only accepts 16 kHz mono and leaves the conversion to the caller).  Rational polyphase FIR on the
 * device: up/down = 16000/rate reduced, 20*max(up,down)+1 Kaiser(beta 5) windowed-sinc taps with unit DC gain — the
 * definition scipy.signal.resample_poly uses.  Channels are averaged.  out_i16 (round to nearest even, saturated) and/or
 * out_f32 (unrounded / 32768) receive *n_out = wdr_resample_n_out(n_frames, rate) samples; WDR_ERR_UNSUPPORTED for rates
 * whose reduced ratio needs more than 40001 taps, WDR_ERR_INVALID if out_cap is too small (*n_out is still set). */
int64_t wdr_resample_n_out(int64_t n_frames, int sample_rate);
int wdr_resample_i16(const int16_t* pcm, int64_t n_frames, int channels, int sample_rate, int16_t* out_i16, float* out_f32,
                     int64_t out_cap, int64_t* n_out);

/* ---- log-mel (whisper.cpp log_mel_spectrogram inside state.full, src/transcribe.rs:389) --------- */
/* Mel front end: holds the [n_mel][201] filterbank (read from the ggml model file upstream), the
 * Hann window and the twiddle tables on the device.  n_mel = 80 or 128. */
typedef struct wdr_mel wdr_mel;
wdr_mel* wdr_mel_init(const float* filters /* [n_mel][201] host */, int n_mel, int device);
void wdr_mel_free(wdr_mel*);
/* n_len whisper.cpp computes for an n-sample buffer: (n + 480000) / 160. */
int wdr_mel_n_len(int n_samples);
/* Whole-buffer log-mel exactly as whisper.cpp lays it out: out[n_mel][n_len], mel-major, normalised
 * with the buffer-global max (clamp to max-8, (x+4)/4) when normalize != 0, raw log10 otherwise.
 * Host pointers.  Returns n_len. */
int wdr_log_mel_f32(wdr_mel*, const float* pcm, int n, int normalize, float* out);
int wdr_log_mel_i16(wdr_mel*, const int16_t* pcm, int n, int normalize, float* out);
/* Batched 30 s windows, the sharded mode of SURVEY §0.4: chunk b = pcm[b*chunk_stride ..][0..n_valid[b])
 * (n_valid NULL => all 480000 valid); produces out[b][n_mel][3000] (the frames the encoder consumes,
 * per-chunk max normalisation == whisper.cpp on that chunk alone) and, if out_max != NULL, the raw
 * per-chunk max.  DEVICE pointers, asynchronous on `stream`. */
int wdr_log_mel_batch_f32_dev(wdr_mel*, const float* pcm, int64_t chunk_stride, const int32_t* n_valid, int n_chunks,
                              int normalize, float* out, float* out_max, void* stream);
int wdr_log_mel_batch_i16_dev(wdr_mel*, const int16_t* pcm, int64_t chunk_stride, const int32_t* n_valid, int n_chunks,
                              int normalize, float* out, float* out_max, void* stream);
/* Same, HOST pointers (H2D + kernel + D2H inside the call; pcm may be pinned or pageable). */
int wdr_log_mel_batch_i16(wdr_mel*, const int16_t* pcm, int64_t chunk_stride, const int32_t* n_valid, int n_chunks,
                          int normalize, float* out);

/* ---- DTW word alignment (whisper.cpp whisper_exp_compute_token_level_timestamps_dtw; enabled by
 *      create_context, src/transcribe.rs:115-136; consumed at src/transcribe.rs:272-282) ---------- */
/* median_filter custom op: w[H][N][M] -> out, odd width <= 31, reflect indexing along M. Host ptrs. */
int wdr_median_filter(const float* w, int H, int N, int M, int width, float* out);
/* Steps 4-6 of SURVEY A.6: alignment-head weights w[H][n_tokens][n_audio] -> normalise over tokens
 * (eps 1e-9) -> median(width) over audio -> mean over heads -> negate -> drop the first sot_len rows and
 * the last row.  out[(n_tokens-sot_len-1)][n_audio]. Host ptrs. */
int wdr_dtw_cost(const float* w, int H, int n_tokens, int n_audio, int sot_len, int width, float* out);
/* dtw_and_backtrace: cost x[N][M] -> monotone path; text_idx/time_idx hold >= N+M entries; *path_len
 * receives the length.  Anti-diagonal wavefront on the device + serial integer backtrace.  Optional
 * cost_out/trace_out ((N+1)*(M+1) each, may be NULL) return the accumulated-cost and trace matrices. Host ptrs. */
int wdr_dtw(const float* x, int N, int M, int32_t* text_idx, int32_t* time_idx, int* path_len,
            float* cost_out, int32_t* trace_out);
/* Batched form over independent windows: x[b] is [N[b]][M[b]] at x + x_offset[b]; outputs at stride
 * max_path per window.  DEVICE pointers for x / outputs; N, M, x_offset are HOST arrays. */
int wdr_dtw_batch_dev(const float* x, const int64_t* x_offset, const int32_t* N, const int32_t* M, int n_windows,
                      int32_t* text_idx, int32_t* time_idx, int32_t* path_len, int max_path, void* stream);

/* ---- Kaldi fbank (knf-rs compute_fbank inside EmbeddingExtractor::compute, src/transcribe.rs:466) */
/* Frames for an n-sample segment with snip_edges: 0 if n < 400 else 1 + (n-400)/160. */
int wdr_fbank_frames(int n_samples);
/* 25 ms / 10 ms povey-window, preemph 0.97, DC removal, 512-pt power spectrum, n_bins HTK-mel bins
 * 20 Hz..Nyquist, log(max(e, FLT_EPSILON)); subtract_mean != 0 applies pyannote-rs' per-column mean
 * subtraction.  pcm is int16 (cast to float WITHOUT scaling, as pyannote-rs does).  out[T][n_bins].
 * Host pointers.  Returns T, or WDR_ERR_TOO_SHORT. */
int wdr_kaldi_fbank_i16(const int16_t* pcm, int n, int n_bins, int subtract_mean, float* out);
/* DEVICE pointers, batched over segments laid out back to back: segment s = pcm[seg_offset[s] ..
 * seg_offset[s+1]); features of segment s start at frame feat_offset[s] of out.  seg_offset /
 * feat_offset are DEVICE int64 arrays of n_segments+1 entries (feat_offset = prefix sum of
 * wdr_fbank_frames).  total_frames = feat_offset[n_segments] (host value). */
int wdr_kaldi_fbank_batch_i16_dev(const int16_t* pcm, const int64_t* seg_offset, const int64_t* feat_offset,
                                  int n_segments, int64_t total_frames, int n_bins, int subtract_mean,
                                  float* out, void* stream);

/* ---- speaker embeddings (pyannote_rs::EmbeddingExtractor, src/transcribe.rs:343, 466-467) ---------------- */
/* WeSpeaker ResNet34 (the north-star's model; the crate downloads the CAM++ export, same "feats" -> "embs" contract): int16
 * samples cast to f32 without scaling -> Kaldi fbank (80 bins) -> per-column mean subtraction -> ResNet34 -> TSTP -> 256-d. */
typedef struct wdr_emb wdr_emb;
wdr_emb* wdr_emb_init(const char* path /* a WeSpeaker ResNet34 .onnx export; NULL: seeded weights */, uint64_t seed, int device);   /* EmbeddingExtractor::new(path) */
void wdr_emb_free(wdr_emb* m);
int wdr_emb_dim(wdr_emb* m);   /* embedding width D: a property of the loaded model (rows of its last Linear): 256 for WeSpeaker ResNet34 */
/* compute(&samples): HOST pointers.  WDR_ERR_TOO_SHORT when the segment yields no fbank frame (< 400 samples): the crate maps
 * that error to speaker "?" (src/transcribe.rs:468-476). */
int wdr_emb_compute_i16(wdr_emb* m, const int16_t* pcm, int64_t n, float* out /* [D] */);
/* All segments of a recording in one call: segment s = pcm[seg_offset[s] .. seg_offset[s+1]) (HOST arrays); out [n][D];
 * status[s] = 0 or WDR_ERR_TOO_SHORT (that row of out is zero).  _dev: pcm / out are DEVICE pointers, offsets / status HOST. */
int wdr_emb_compute_batch_i16(wdr_emb* m, const int16_t* pcm, const int64_t* seg_offset, int n_segments, float* out, int32_t* status);
int wdr_emb_compute_batch_i16_dev(wdr_emb* m, const int16_t* pcm_dev, const int64_t* seg_offset_host, int n_segments, float* out_dev,
                                  int32_t* status_host, void* stream);
double wdr_emb_last_flops(wdr_emb* m);  /* algorithmic conv FLOPs (2*M*N*K) of the last compute call */
/* Measurement aid: with profiling on, every tcgen05 GEMM and every im2col gather of a compute call is bracketed by a CUDA-event pair on
 * the launching stream; wdr_emb_last_kernel_ms returns their summed device times for the last call (bench.py quotes the GEMMs alone
 * against the tensor roofline). */
int wdr_emb_profile(wdr_emb* m, int enable);
int wdr_emb_last_kernel_ms(wdr_emb* m, double* gemm_ms, double* gather_ms);

/* ---- get_signal_energy (whisper.cpp, used by the token-timestamp heuristic) --------- */
int wdr_signal_energy(const float* pcm, int n, int half_window, float* out);

/* ---- context / state (whisper_context, whisper_state) ------------------------------------------- */
typedef struct wdr_context wdr_context;
typedef struct wdr_state wdr_state;
/* == whisper_alignment_heads_preset; the crate maps model names to these (src/transcribe.rs:117-129). */
enum wdr_aheads_preset {
    WDR_AHEADS_NONE = -1, WDR_AHEADS_TINY_EN = 0, WDR_AHEADS_TINY, WDR_AHEADS_BASE_EN, WDR_AHEADS_BASE, WDR_AHEADS_SMALL_EN,
    WDR_AHEADS_SMALL, WDR_AHEADS_MEDIUM_EN, WDR_AHEADS_MEDIUM, WDR_AHEADS_LARGE_V3, WDR_AHEADS_LARGE_V3_TURBO
};
/* == whisper_context_params (WhisperContextParameters, src/transcribe.rs:102-136), plus the seeded-weights
 * extension (no model file can exist here): arch_name selects the architecture, seed the weights. */
typedef struct wdr_context_params {
    int use_gpu;               /* false is refused: there is no CPU path */
    int gpu_device;
    int flash_attn;            /* accepted; the encoder attention is always the fused tcgen05 kernel */
    int dtw_token_timestamps;
    int dtw_aheads_preset;     /* enum wdr_aheads_preset */
    size_t dtw_mem_size;       /* accepted (src/utils.rs:3-49); workspaces are sized from the batch instead */
    const char* arch_name;     /* "tiny.en" ... "large-v3-turbo" */
    uint64_t seed;
} wdr_context_params;
typedef struct wdr_model_dims {
    int n_audio_state, n_audio_head, n_audio_layer, n_text_layer, n_mels, n_vocab, n_audio_ctx, n_text_ctx, is_multilingual;
    int64_t weight_bytes;
} wdr_model_dims;
wdr_context_params wdr_context_default_params(void);                                         /* whisper_context_default_params */
/* whisper_init_from_file_with_params (WhisperContext::new_with_params, src/transcribe.rs:154).  path: a whisper.cpp
 * ggml-<model>.bin checkpoint (f32 / f16 tensors; names at src/model_manager.rs:162) — geometry, mel filterbank, token strings and
 * weights come from the file, params.dtw_aheads_preset selects the alignment heads as the crate does by model name.  path == NULL:
 * seeded synthetic weights of params.arch_name (no checkpoint exists offline).  Returns NULL on failure, never aborts. */
wdr_context* wdr_init_from_file_with_params(const char* path, wdr_context_params params);
void wdr_free(wdr_context* ctx);                                                             /* whisper_free */
/* Header of a ggml-<model>.bin checkpoint without touching the GPU: hparams[11] = n_vocab, n_audio_ctx, n_audio_state,
 * n_audio_head, n_audio_layer, n_text_ctx, n_text_state, n_text_head, n_text_layer, n_mels, ftype; tensor and token counts.
 * (`path` in wdr_init_from_file_with_params: f32 / f16 checkpoints of the OpenAI geometries; quantised files are refused.) */
int wdr_ggml_probe(const char* path, int32_t* hparams, int32_t* n_tensors, int32_t* n_tokens);
/* The other model files the crate opens, read without touching the GPU (dependency-free readers, csrc/onnx_file.cu, csrc/ggml_file.cu):
 * wdr_onnx_probe: segmentation-3.0.onnx (kind 0, src/engine.rs:90) or the WeSpeaker ResNet34 export (kind 1, src/engine.rs:91) —
 *   info[6] = nodes, constant tensors, graph inputs, graph outputs, parameters the loader takes, embedding width (kind 1).
 *   Fails with the reason when the operator sequence is not that architecture (e.g. the CAM++ export).
 * wdr_onnx_read_param: one extracted parameter under its PyTorch-style name and layout ("lstm.weight_ih_l0_reverse", "layer2.0.conv1.weight",
 *   ...; LSTM gates re-ordered from ONNX i,o,f,c to i,f,g,o; BatchNorm folded); returns the element count (out may be NULL).
 * wdr_silero_probe: ggml-silero-v5.1.2.bin (src/model_manager.rs:305-315) — hparams[20] = version[3], n_encoder_layers,
 *   {in, out, kernel} x 4, lstm_input, lstm_hidden, final_conv_in, final_conv_out. */
int wdr_onnx_probe(const char* path, int kind, int32_t* info);
int64_t wdr_onnx_read_param(const char* path, int kind, const char* name, float* out, int64_t cap);
int wdr_silero_probe(const char* path, int32_t* hparams, int32_t* n_tensors);
int wdr_model_info(const wdr_context* ctx, wdr_model_dims* out);                             /* whisper_model_n_* getters */
wdr_state* wdr_init_state(wdr_context* ctx);                 /* whisper_init_state (ctx.create_state(), src/transcribe.rs:335) */
void wdr_free_state(wdr_state* state);                       /* whisper_free_state */
/* librosa slaney filterbank [n_mel][201] that whisper.cpp reads from the model file. Host pointer. */
int wdr_mel_filters(int n_mel, float* out);

/* ---- encoder (whisper_encode / the encoder half of whisper_full_with_state, src/transcribe.rs:389) --- */
/* whisper_encode semantics: normalised mel[n_mel][n_len] (host), window at frame mel_offset, zero-extended to
 * 3000 frames -> hidden[1500][n_audio_state] fp32 (host). */
int wdr_encode(wdr_context* ctx, wdr_state* state, const float* mel, int n_len, int mel_offset, float* out_hidden);
/* Sharded mode: B independent 30 s windows of int16 PCM -> log-mel -> encoder -> hidden[B][1500][d].
 * _dev: DEVICE pointers, asynchronous on stream.  Without suffix: HOST pointers, copies inside the call. */
int wdr_encode_chunks_i16_dev(wdr_context* ctx, wdr_state* state, const int16_t* pcm, int64_t chunk_stride, const int32_t* n_valid,
                              int n_chunks, float* out_hidden, void* stream);
int wdr_encode_chunks_i16(wdr_context* ctx, wdr_state* state, const int16_t* pcm, int64_t chunk_stride, const int32_t* n_valid,
                          int n_chunks, float* out_hidden /* may be NULL: the result stays in the state, as whisper_encode leaves it */);
/* Per-window mean |hidden| of the encoder output held in the state (n <= windows of the last encode call): a
 * 4-byte-per-window device->host read that proves the encode finished without moving the hidden states. */
int wdr_state_hidden_digest(wdr_state* state, float* out, int n);
/* Per-kernel-class CUDA-event timing on the launching stream.  Classes: 0 mel, 1 mel re-layout, 2 tcgen05 GEMM (encoder + cross-KV),
 * 3 encoder attention, 4 layernorm, 5 decoder small kernels (embed / LN / self-attention / GELU / sampler), 6 dtw, 7 other,
 * 8 decoder cross-attention (HBM-bound), 9 decoder weight-streaming GEMMs, 10 batched cross-attention of the DTW pass.  collect() sums the finished records (caller has
 * synchronised) into ms[] / launches[] (>= 11 entries each) and returns the number of classes. */
int wdr_profile_enable(wdr_state* state, int enable);
int wdr_profile_collect(wdr_state* state, double* ms, int32_t* launches, int n_classes);
/* B200 extension (no whisper.h equivalent): how many lanes wdr_full_batch_* cuts its windows into.  Each lane is a host thread
 * with its own streams and workspaces, so one lane's latency-bound decode chain overlaps another's HBM-bound cross-attention and
 * tensor-bound encoder.  0 = default (environment WDR_LANES, else 1); results do not depend on the lane count. */
int wdr_state_set_lanes(wdr_state* state, int n_lanes);
/* Encoder self-attention alone: qk bf16 [B*T][2d] (query | key), vt bf16 [d][ldt] = V transposed, window b's tokens at
 * columns b*round_up(T,8) + t (pad columns zero) -> out bf16 [B*T][d].  DEVICE pointers. */
int wdr_encoder_attention_dev(const uint16_t* qk, const uint16_t* vt, int64_t ldt, int n_chunks, int T, int n_head, int d_model,
                              uint16_t* out, void* stream);

/* ---- full transcription (whisper_full_with_state: state.full, src/transcribe.rs:389; results :393-412, :252-282) -- */
/* == whisper_token_data (WhisperToken::token_data(), src/transcribe.rs:272-282). t0/t1/t_dtw in centiseconds relative
 * to the start of the buffer passed to the call; t_dtw = -1 when DTW did not reach the token. */
typedef struct wdr_token_data {
    int32_t id, tid;
    float p, plog, pt, ptsum;
    int64_t t0, t1, t_dtw;
    float vlen;
} wdr_token_data;
enum wdr_sampling_strategy { WDR_SAMPLING_GREEDY = 0, WDR_SAMPLING_BEAM_SEARCH = 1 };
/* == whisper_full_params, the fields the crate sets (setup_params, src/transcribe.rs:20-87) plus whisper.cpp's defaults
 * for the rest (wdr_full_default_params == whisper_full_default_params, temperature_inc = 0.2 included).  Supported decoding:
 * greedy and beam search (beam_size <= 8) with whisper_full's temperature ladder `for (t = temperature; t < 1 + 1e-6;
 * t += temperature_inc)`: windows whose decode fails whisper_full's success test are decoded again at the next temperature;
 * `temperature` may start above 0 (the crate forwards advanced.temperature, src/transcribe.rs:58-68).  Above temperature 0
 * both strategies run max(1, greedy.best_of) (<= 8) decoders per window; the greedy strategy draws every token from
 * std::discrete_distribution, decoder j of a window seeded std::mt19937(j) (upstream's stream runs on across the windows of a
 * call; here every window restarts it, so windows stay independent and shardable).  no_speech_prob is always taken from the
 * raw logits of the prompt decode (before the temperature division), as upstream does.
 * language given or "auto"; single_segment = 1 as the crate always sets (src/transcribe.rs:46).  Parameters that are not
 * implemented (offset_ms, duration_ms, max_len, max_tokens, audio_ctx, suppress_nst, no_timestamps, single_segment = 0) are
 * refused with WDR_ERR_UNSUPPORTED, never silently ignored.  Strings are borrowed for the call. */
typedef struct wdr_full_params {
    int strategy;            /* enum wdr_sampling_strategy */
    int n_threads;           /* accepted, unused */
    int n_max_text_ctx;
    int offset_ms, duration_ms;
    int translate, no_context, no_timestamps, single_segment;
    int print_special, print_progress, print_realtime, print_timestamps;
    int token_timestamps;    /* heuristic t0/t1 per token */
    float thold_pt, thold_ptsum;
    int max_len, split_on_word, max_tokens;
    int audio_ctx;
    const char* initial_prompt;      /* tokenised with the context's vocabulary (wdr_tokenize); then replaces prompt_tokens, as upstream */
    const int32_t* prompt_tokens;    /* ids appended to the text context ([PREV] + tokens + sot sequence) */
    int prompt_n_tokens;
    const char* language;            /* "en", "de", ...; NULL = "en" */
    int detect_language;
    int suppress_blank, suppress_nst;
    float temperature, max_initial_ts, length_penalty;
    float temperature_inc, entropy_thold, logprob_thold, no_speech_thold;
    int greedy_best_of, beam_size;
    float beam_patience;
    void (*progress_callback)(wdr_context* ctx, wdr_state* state, int progress, void* user_data);
    void* progress_callback_user_data;
    bool (*abort_callback)(void* user_data);   /* polled between decoder steps; true -> WDR_ERR_ABORTED (src/transcribe.rs:348-350) */
    void* abort_callback_user_data;
} wdr_full_params;
wdr_full_params wdr_full_default_params(int strategy);                                     /* whisper_full_default_params */
/* state.full(params, &samples) on one buffer of any length (what the crate submits per SpeechSegment, src/transcribe.rs:376-389).
 * n <= 480000: one window.  Longer: whisper_full's sequential seek loop (global mel max, window k+1 starts where window k's last
 * timestamp says and is prompted with [PREV] + prompt_past; timestamp state carried) — replicas only, it does not shard.
 * 0 = ok; results stay in the state until the next call.  Host pointers. */
int wdr_full_with_state(wdr_context* ctx, wdr_state* state, wdr_full_params params, const float* pcm, int n);
int wdr_full_with_state_i16(wdr_context* ctx, wdr_state* state, wdr_full_params params, const int16_t* pcm, int n);
/* Sharded mode: n_chunks independent buffers of <= 30 s each (chunk b = pcm[b*chunk_stride ..][0..n_valid[b]),
 * n_valid NULL = 480000), each treated exactly as its own wdr_full_with_state call; mel + encoder + cross-KV + greedy decode +
 * token timestamps + DTW run batched on the device.  Results: segments of all chunks in chunk order. */
int wdr_full_batch_i16(wdr_context* ctx, wdr_state* state, wdr_full_params params, const int16_t* pcm, int64_t chunk_stride,
                       const int32_t* n_valid, int n_chunks);
/* Same with the PCM already resident in HBM (DEVICE pointer, chunk_stride >= 480000 elements; n_valid stays a HOST array):
 * the arm bench.py times as `value`.  Results are still gathered into the state (host) before the call returns. */
int wdr_full_batch_i16_dev(wdr_context* ctx, wdr_state* state, wdr_full_params params, const int16_t* pcm_dev, int64_t chunk_stride,
                           const int32_t* n_valid, int n_chunks);
int wdr_full_n_segments_from_state(wdr_state* state);                                      /* state.full_n_segments(), :397 */
int wdr_full_get_segment_chunk_from_state(wdr_state* state, int i_segment);                /* chunk the segment belongs to (batch calls) */
int64_t wdr_full_get_segment_t0_from_state(wdr_state* state, int i_segment);               /* start_timestamp(), cs */
int64_t wdr_full_get_segment_t1_from_state(wdr_state* state, int i_segment);               /* end_timestamp(), cs */
const char* wdr_full_get_segment_text_from_state(wdr_state* state, int i_segment);         /* to_str(); owned by the state */
float wdr_full_get_segment_no_speech_prob_from_state(wdr_state* state, int i_segment);
int wdr_full_n_tokens_from_state(wdr_state* state, int i_segment);                         /* n_tokens(), :252 */
int32_t wdr_full_get_token_id_from_state(wdr_state* state, int i_segment, int i_token);
const char* wdr_full_get_token_text_from_state(wdr_context* ctx, wdr_state* state, int i_segment, int i_token);  /* to_str_lossy(), :257 */
wdr_token_data wdr_full_get_token_data_from_state(wdr_state* state, int i_segment, int i_token);                /* token_data(), :272 */
int wdr_full_lang_id_from_state(wdr_state* state);                                         /* full_lang_id_from_state(), :393 */
/* language = "auto" (or detect_language) on a multilingual model: whisper_lang_auto_detect — [SOT] decoded against the buffer's
 * first window, arg-max over the language tokens.  In a batch call every 30 s buffer is detected on its own (each is its own
 * state.full); this returns buffer i's language, wdr_full_lang_id_from_state the first one's. */
int wdr_full_get_chunk_lang_id_from_state(wdr_state* state, int i_chunk);
const char* wdr_lang_str(int id);                                                          /* whisper_rs::get_lang_str, :394 */
int wdr_lang_id(const char* lang);                                                         /* whisper_lang_id */
const char* wdr_token_to_str(wdr_context* ctx, int32_t token);                             /* whisper_token_to_str */
/* whisper_tokenize (what whisper_full applies to `initial_prompt`; the crate feeds the previous segment's text back through it,
 * src/transcribe.rs:383-386): GPT-2 style word split + longest vocabulary entry at every position.  Returns the number of tokens
 * written, or the negated count if it exceeds n_max_tokens.  Host code (no device work).  wdr_tokenize_with_vocab runs the same
 * tokenizer over an explicit vocabulary (token_strings[i] = text of id i; NULL entries are skipped). */
int wdr_tokenize(wdr_context* ctx, const char* text, int32_t* tokens, int n_max_tokens);
int wdr_tokenize_with_vocab(const char* const* token_strings, int n_tokens, const char* text, int32_t* tokens, int n_max_tokens);
/* Per-chunk decoder summary of the last full call: info[8] = {seek_delta, failed, completed, n_sampled, has_ts, result_len,
 * seek_end, n_segments}; *no_speech_prob optional.  For parity tests. */
/* Device time (CUDA events on the compute stream) of the phases of the last full call, summed over its groups (and lanes):
 * ms[5] = log-mel + encoder | cross-KV projection | greedy decode loop | batched DTW pass | DTW cost + wavefront + backtrace;
 * decode_steps = greedy iterations run. */
int wdr_full_get_phase_ms(wdr_state* state, double* ms, int32_t* decode_steps);
int wdr_full_get_chunk_info_from_state(wdr_state* state, int i_chunk, int32_t* info, float* no_speech_prob);
/* dec_cross_attn_kernel of the last full call, counted on the device: launches, and (launch, window) pairs in which the window was
 * still decoding and really streamed its K_c / V_c (finished windows exit at once).  bench.py's roofline uses live_windows, so the
 * algorithmic bytes follow the windows that were live (with checkpoints that emit EOT a launch serves fewer than B windows). */
int wdr_full_get_cross_attn_stats(wdr_state* state, int64_t* launches, int64_t* live_windows);
/* Temperature of the ladder (temperature, +temperature_inc, ... <= 1) whose result stands for chunk i of the last full call;
 * -1 if i is out of range. */
float wdr_full_get_chunk_temperature_from_state(wdr_state* state, int i_chunk);
/* The host-side sampler of the temperature ladder for the greedy strategy (whisper_sample_token, best = false): n_draws
 * consecutive draws from std::discrete_distribution over expf(logprobs) (-inf = probability 0) with one std::mt19937(seed).
 * Pure host code (no device needed); exposed so that the draw can be checked bit for bit against the oracle's restatement. */
int wdr_sample_discrete(const float* logprobs, int n, uint32_t seed, int n_draws, int32_t* ids);
/* Stage-level decoder access for parity tests: teacher-forced pass of `n_seq` tokens over window 0.. of the last encode/full call.
 * enc: optional HOST encoder output [n_chunks][1500][d] to install first (NULL = keep the state's).  seq: HOST [n_chunks][n_seq].
 * logits_out: HOST [n_chunks][n_seq][n_vocab] or NULL.  aheads_out: HOST [n_chunks][n_aheads][n_seq][1500] or NULL (needs a DTW
 * context). */
int wdr_decode_teacher_forced(wdr_context* ctx, wdr_state* state, const float* enc, int n_chunks, const int32_t* seq, int n_seq,
                              float* logits_out, float* aheads_out);

/* ---- Silero VAD (WhisperVadContext, reference src/vad.rs:15-43; whisper.cpp whisper_vad_*) ----------------- */
typedef struct wdr_vad wdr_vad;
typedef struct wdr_vad_segments wdr_vad_segments;
typedef struct wdr_vad_context_params { int n_threads, use_gpu, gpu_device; uint64_t seed; } wdr_vad_context_params;  /* + seed: synthetic weights */
/* == whisper_vad_params (WhisperVadParams; the crate sets min_silence_duration_ms = 100, src/vad.rs:22) */
typedef struct wdr_vad_params {
    float threshold; int min_speech_duration_ms, min_silence_duration_ms; float max_speech_duration_s; int speech_pad_ms; float samples_overlap;
} wdr_vad_params;
wdr_vad_context_params wdr_vad_default_context_params(void);                     /* whisper_vad_default_context_params */
wdr_vad_params wdr_vad_default_params(void);                                      /* whisper_vad_default_params */
wdr_vad* wdr_vad_init_from_file_with_params(const char* path, wdr_vad_context_params params);  /* path: ggml-silero-v5.1.2.bin; NULL: seeded weights */
void wdr_vad_free(wdr_vad* v);
/* whisper_vad_detect_speech: one probability per 512-sample frame (last frame zero padded); LSTM state reset per call. Host ptr. */
int wdr_vad_detect_speech(wdr_vad* v, const float* pcm, int n);
int wdr_vad_n_probs(wdr_vad* v);
const float* wdr_vad_probs(wdr_vad* v);
/* Batched form over independent streams (files / shards) of int16 PCM: stream s = pcm[offsets[s] .. +n_samples[s]); probabilities of
 * stream s land at probs_out[frame_offsets_out[s] ..).  frame_offsets_out has n_streams+1 entries.  Host pointers. */
int wdr_vad_detect_speech_batch_i16(wdr_vad* v, const int16_t* pcm, const int64_t* offsets, const int32_t* n_samples, int n_streams,
                                    float* probs_out, int64_t* frame_offsets_out);
wdr_vad_segments* wdr_vad_segments_from_probs(wdr_vad* v, wdr_vad_params params);                         /* whisper_vad_segments_from_probs */
wdr_vad_segments* wdr_vad_segments_from_probs_array(const float* probs, int n_probs, wdr_vad_params params);  /* same on a caller-supplied array (host logic only) */
wdr_vad_segments* wdr_vad_segments_from_samples(wdr_vad* v, wdr_vad_params params, const float* pcm, int n);  /* vad.segments_from_samples, src/vad.rs:31 */
int wdr_vad_segments_n(wdr_vad_segments* s);
float wdr_vad_segments_get_segment_t0(wdr_vad_segments* s, int i);   /* centiseconds, as src/vad.rs:40-43 consumes them */
float wdr_vad_segments_get_segment_t1(wdr_vad_segments* s, int i);
void wdr_vad_free_segments(wdr_vad_segments* s);

/* ---- pyannote segmentation (pyannote_rs::get_segments, reference src/engine.rs:117-122) ------------------------- */
typedef struct wdr_seg wdr_seg;
typedef struct wdr_seg_result wdr_seg_result;
wdr_seg* wdr_seg_init(const char* path /* segmentation-3.0.onnx; NULL: seeded weights */, uint64_t seed, int device);
void wdr_seg_free(wdr_seg* m);
int wdr_seg_n_windows(int64_t n_samples);                              /* n / 160000 + 1: pyannote-rs pads `window - len % window` zeros, so an exact
                                                                          multiple of 10 s gets one more (silent) window */
/* PyanNet over every 10 s window (raw int16 values as f32, zero padded): scores[n_windows][589][7] log-probabilities. Host ptrs.
 * Returns n_windows. */
int wdr_seg_scores_i16(wdr_seg* m, const int16_t* pcm, int64_t n, float* scores);
/* get_segments: windows -> scores -> per-frame argmax != 0 state machine (frame 270 samples, first frame at sample 721). */
wdr_seg_result* wdr_seg_get_segments(wdr_seg* m, const int16_t* pcm, int64_t n);
/* The state machine alone on caller-supplied scores (host logic; bit-exact given the scores).  n_samples = the ORIGINAL sample
 * count: sample ranges are clamped to it (start to n - 1, end to n) as pyannote-rs does, so no segment carries padding zeros. */
wdr_seg_result* wdr_seg_segments_from_scores(const float* scores, int n_windows, int64_t n_samples);
int wdr_seg_result_n(wdr_seg_result* r);
double wdr_seg_result_start(wdr_seg_result* r, int i);                 /* Segment.start, seconds (f64) */
double wdr_seg_result_end(wdr_seg_result* r, int i);
int64_t wdr_seg_result_sample_range(wdr_seg_result* r, int i, int64_t* i1);   /* [i0, i1) into the input (clamped to its length) */
const int16_t* wdr_seg_result_samples(wdr_seg_result* r, int i, int64_t* count);  /* Segment.samples (owned by the result) */
void wdr_seg_result_free(wdr_seg_result* r);

/* ---- speaker assignment (pyannote_rs::EmbeddingManager, src/transcribe.rs:342, 480-492) -------------------- */
typedef struct wdr_spk wdr_spk;
wdr_spk* wdr_spk_init(size_t max_speakers);            /* EmbeddingManager::new(max_speakers); SIZE_MAX = unlimited (src/engine.rs:108-111) */
void wdr_spk_free(wdr_spk* m);
int wdr_spk_count(wdr_spk* m);                         /* get_all_speakers().len() */
/* search_speaker(embedding, threshold): best cosine match if sim > threshold (strict), else a new id while below the cap.
 * Returns the speaker id (>= 1) or 0 for None (the crate renders "?", src/transcribe.rs:493-495). Host pointer. */
int wdr_spk_search(wdr_spk* m, const float* emb, int dim, float threshold);
/* get_best_speaker_match(embedding): best cosine match regardless of threshold; WDR_ERR_INVALID when no speaker is stored. */
int wdr_spk_best_match(wdr_spk* m, const float* emb, int dim);
/* The crate's per-segment policy (cap reached -> get_best_speaker_match, else search_speaker; src/transcribe.rs:480-492) applied to n
 * embeddings in order; labels[i] = id >= 1 or 0 for "?".  Returns the number of speakers.  Host logic (what a multi-GPU host runs on the
 * all-gathered table). */
int wdr_spk_assign_batch(wdr_spk* m, const float* emb, int n, int dim, float threshold, int32_t* labels);
/* Pairwise cosine similarity S[N][N] of embeddings emb[N][D] on the device (host pointers). */
int wdr_cosine_matrix(const float* emb, int N, int D, float* S);
/* The crate's per-segment policy (cap reached -> best match, else search/create) as a pure function of S, segments in time
 * order; labels[i] = speaker id >= 1 or 0 for "?".  Returns the number of speakers.  Bit-exact given S. */
int wdr_cluster_leader(const float* S, int N, float threshold, size_t max_speakers, int32_t* labels);
/* Average-linkage agglomerative clustering of S on the device: merges the first maximal pair while its similarity > threshold.
 * labels[i] = 1.. in order of each cluster's smallest member.  Returns the number of clusters.  Bit-exact given S. */
int wdr_cluster_agglomerative(const float* S, int N, float threshold, int32_t* labels);

/* ---- multi-GPU exchange ----------------------------------------------------------------------------------------------
 * The path shards by independent unit (30 s window, 10 s diarization window, speech segment): one context per GPU, static contiguous
 * blocks of units per rank, weights replicated, NO data-path collective.  Its one exchange is the all-gather of per-rank speaker
 * embeddings ahead of global clustering: NCCL (opened at run time: libnccl.so.2 or $WDR_NCCL_LIB) over NVLink / NVSwitch on device
 * buffers.  The reference itself is single-GPU (`gpu_device`, src/engine.rs:14, src/transcribe.rs:110-112); a host that drives N GPUs
 * creates one wdr_dist per rank: N processes (rank 0 calls wdr_dist_get_unique_id, the id travels over the host's own channel, every rank
 * calls wdr_dist_init) or N threads of one process (the same, or wdr_dist_init_all).  n_ranks = 1 needs no NCCL. */
#define WDR_DIST_ID_BYTES 128
typedef struct wdr_dist wdr_dist;
int wdr_dist_available(void);                                   /* 1 if NCCL could be opened */
int wdr_dist_nccl_version(void);                                /* ncclGetVersion, 0 if unavailable */
int wdr_dist_get_unique_id(uint8_t* id /* [WDR_DIST_ID_BYTES] */);
wdr_dist* wdr_dist_init(const uint8_t* id, int n_ranks, int rank, int device);     /* collective: every rank calls it; NULL on failure */
int wdr_dist_init_all(int n_gpus, const int* devices /* NULL: 0 .. n_gpus-1 */, wdr_dist** out /* [n_gpus] */);
void wdr_dist_free(wdr_dist* d);
int wdr_dist_size(wdr_dist* d);
int wdr_dist_rank(wdr_dist* d);
/* emb [n_local][D] of this rank -> out [sum_r n_r][D] in rank order on EVERY rank (segments stay in time order when shards are
 * contiguous); counts_out[r] = n_r (HOST, may be NULL).  Ranks may hold different, also zero, row counts.  normalize != 0: rows are
 * scaled to unit L2 norm on the way into the send buffer (the cosine matrix of the gathered table is then a plain E E^T).  Returns the
 * number of rows gathered.  _dev: emb / out are DEVICE pointers (out_cap_rows rows available); the collective runs on `stream`
 * (NULL = the default stream) after whatever the caller queued there, and the call returns once it has completed; without suffix:
 * HOST pointers. */
int wdr_allgather_embeddings_dev(wdr_dist* d, const float* emb_dev, int n_local, int D, int normalize, float* out_dev, int64_t out_cap_rows,
                                 int32_t* counts_out, void* stream);
int wdr_allgather_embeddings(wdr_dist* d, const float* emb, int n_local, int D, int normalize, float* out, int64_t out_cap_rows,
                             int32_t* counts_out);

/* Bring-up aid: copies a decoder workspace buffer of the last step to the host as fp32 (0 x, 1 h, 2 att, 3 ff, 4 layer-0 cross K|V,
 * 5 split-K partials). */
int wdr_debug_decoder_read(wdr_state* state, int which, float* out, int64_t count);

/* ---- stage-level kernels (not in whisper.h; for parity tests and roofline measurement) -------------- */
/* tcgen05 bf16 GEMM: D[M][N] = A[M][K] * W[N][K]^T with a fused epilogue (the ggml mul_mat / ONNX Gemm of the
 * encoder, decoder and embedding nets).  bf16 values are passed as uint16_t.  A rows are n_batch groups of
 * rows_per_batch rows (row stride lda, batch stride a_batch_stride, elements).  kb_per_tap > 0 selects the
 * implicit-GEMM (conv1d) addressing: K is cut into taps of kb_per_tap*64 columns, tap t reads A rows r + t.
 * epilogue: 0 bias->bf16, 1 bias+GELU->bf16, 2 resid+bias->f32 (resid_or_pos = resid[M][ldc]),
 * 3 GELU(bias)+pos->f32 (resid_or_pos = pos[rows_per_batch][N]), 4 QKV (columns >= n_split stored transposed
 * into out_t[n - n_split][batch * t_batch_stride + row_in_batch], row stride ldt; t_batch_stride 0 = rows_per_batch),
 * 5 bias->f32, 9 bias->bf16 stored head-major (the decoder's cross K|V cache: n_split = rows per group g, row = g*n_split + t,
 * n = s*(N/2) + h*64 + c -> out[((g*H + h)*2 + s)*n_split + t][c]).  DEVICE pointers, asynchronous on stream. */
int wdr_gemm_bf16_dev(const uint16_t* A, int64_t lda, int rows_per_batch, int n_batch, int64_t a_batch_stride,
                      const uint16_t* W, int64_t ldw, int N, int K, int kb_per_tap, int a_cols, const float* bias,
                      int epilogue, void* out, int64_t ldc, const float* resid_or_pos, uint16_t* out_t, int64_t ldt,
                      int n_split, int64_t t_batch_stride, void* stream);

/* N-tile width (64 / 128 / 256) the last wdr_gemm_bf16_dev call on this thread ran with: parity tests assert that the shapes of the
 * benchmark configuration really select the 128 x 256 tile. */
int wdr_gemm_last_tile_n(void);

#ifdef __cplusplus
}
#endif
#endif /* WDR_H */
