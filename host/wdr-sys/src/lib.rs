//! Raw bindings to `include/wdr.h`, one declaration per C entry point, each citing the reference call site it replaces
//! (paths relative to tmoroney/whisper-diarize-rs).  Hand-maintained (no bindgen in the image); layouts are checked by the
//! `size_of` assertions at the bottom against the values `tests/test_capi_cpu.py` reads from the C side.
#![allow(non_camel_case_types)]
use libc::{c_char, c_double, c_float, c_int, c_void, size_t};

pub const WDR_OK: c_int = 0;
pub const WDR_ERR_NO_DEVICE: c_int = -2;
pub const WDR_ERR_TOO_SHORT: c_int = -6; // pyannote-rs: "fbank yields 0 frames" -> speaker "?" (src/transcribe.rs:468-476)
pub const WDR_SAMPLING_GREEDY: c_int = 0;
pub const WDR_SAMPLING_BEAM_SEARCH: c_int = 1;

#[repr(C)] pub struct wdr_context { _p: [u8; 0] }
#[repr(C)] pub struct wdr_state { _p: [u8; 0] }
#[repr(C)] pub struct wdr_vad { _p: [u8; 0] }
#[repr(C)] pub struct wdr_vad_segments { _p: [u8; 0] }
#[repr(C)] pub struct wdr_seg { _p: [u8; 0] }
#[repr(C)] pub struct wdr_seg_result { _p: [u8; 0] }
#[repr(C)] pub struct wdr_emb { _p: [u8; 0] }
#[repr(C)] pub struct wdr_spk { _p: [u8; 0] }
#[repr(C)] pub struct wdr_dist { _p: [u8; 0] }
pub const WDR_DIST_ID_BYTES: usize = 128;

/// == whisper_context_params (WhisperContextParameters, src/transcribe.rs:102-136) + the synthetic-weights extension.
#[repr(C)] #[derive(Clone, Copy)]
pub struct wdr_context_params {
    pub use_gpu: c_int, pub gpu_device: c_int, pub flash_attn: c_int,
    pub dtw_token_timestamps: c_int, pub dtw_aheads_preset: c_int, pub dtw_mem_size: size_t,
    pub arch_name: *const c_char, pub seed: u64,
}

/// == whisper_token_data (WhisperToken::token_data(), src/transcribe.rs:272-282).
#[repr(C)] #[derive(Clone, Copy, Debug)]
pub struct wdr_token_data {
    pub id: i32, pub tid: i32, pub p: c_float, pub plog: c_float, pub pt: c_float, pub ptsum: c_float,
    pub t0: i64, pub t1: i64, pub t_dtw: i64, pub vlen: c_float,
}

pub type wdr_abort_callback = Option<unsafe extern "C" fn(user_data: *mut c_void) -> bool>;
pub type wdr_progress_callback = Option<unsafe extern "C" fn(ctx: *mut wdr_context, state: *mut wdr_state, progress: c_int, user_data: *mut c_void)>;

/// == whisper_full_params as setup_params fills it (src/transcribe.rs:20-87).  Field order follows include/wdr.h.
#[repr(C)] #[derive(Clone, Copy)]
pub struct wdr_full_params {
    pub strategy: c_int, pub n_threads: c_int, pub n_max_text_ctx: c_int, pub offset_ms: c_int, pub duration_ms: c_int,
    pub translate: c_int, pub no_context: c_int, pub no_timestamps: c_int, pub single_segment: c_int,
    pub print_special: c_int, pub print_progress: c_int, pub print_realtime: c_int, pub print_timestamps: c_int,
    pub token_timestamps: c_int, pub thold_pt: c_float, pub thold_ptsum: c_float, pub max_len: c_int, pub split_on_word: c_int, pub max_tokens: c_int,
    pub audio_ctx: c_int, pub initial_prompt: *const c_char, pub prompt_tokens: *const i32, pub prompt_n_tokens: c_int,
    pub language: *const c_char, pub detect_language: c_int, pub suppress_blank: c_int, pub suppress_nst: c_int,
    pub temperature: c_float, pub max_initial_ts: c_float, pub length_penalty: c_float, pub temperature_inc: c_float,
    pub entropy_thold: c_float, pub logprob_thold: c_float, pub no_speech_thold: c_float,
    pub greedy_best_of: c_int, pub beam_size: c_int, pub beam_patience: c_float,
    pub progress_callback: wdr_progress_callback, pub progress_callback_user_data: *mut c_void,
    pub abort_callback: wdr_abort_callback, pub abort_callback_user_data: *mut c_void,
}

#[repr(C)] #[derive(Clone, Copy)]
pub struct wdr_vad_context_params { pub n_threads: c_int, pub use_gpu: c_int, pub gpu_device: c_int, pub seed: u64 }
/// == whisper_vad_params (WhisperVadParams, src/vad.rs:21-22).
#[repr(C)] #[derive(Clone, Copy)]
pub struct wdr_vad_params {
    pub threshold: c_float, pub min_speech_duration_ms: c_int, pub min_silence_duration_ms: c_int,
    pub max_speech_duration_s: c_float, pub speech_pad_ms: c_int, pub samples_overlap: c_float,
}

extern "C" {
    pub fn wdr_last_error() -> *const c_char;
    pub fn wdr_device_count() -> c_int;
    // ---- whisper-rs surface -------------------------------------------------------------------------------------
    pub fn wdr_context_default_params() -> wdr_context_params;                                    // WhisperContextParameters::default(), src/transcribe.rs:102
    pub fn wdr_init_from_file_with_params(path: *const c_char, p: wdr_context_params) -> *mut wdr_context; // WhisperContext::new_with_params, :154 (NULL on failure, never aborts)
    pub fn wdr_free(ctx: *mut wdr_context);
    pub fn wdr_init_state(ctx: *mut wdr_context) -> *mut wdr_state;                               // ctx.create_state(), :335
    pub fn wdr_free_state(st: *mut wdr_state);
    pub fn wdr_full_default_params(strategy: c_int) -> wdr_full_params;                           // FullParams::new, :37
    pub fn wdr_full_with_state(ctx: *mut wdr_context, st: *mut wdr_state, p: wdr_full_params, pcm: *const c_float, n: c_int) -> c_int; // state.full, :389
    pub fn wdr_full_with_state_i16(ctx: *mut wdr_context, st: *mut wdr_state, p: wdr_full_params, pcm: *const i16, n: c_int) -> c_int; // :380-389 without the f32 copy
    pub fn wdr_full_batch_i16(ctx: *mut wdr_context, st: *mut wdr_state, p: wdr_full_params, pcm: *const i16, chunk_stride: i64,
                              n_valid: *const i32, n_chunks: c_int) -> c_int;                     // B200 extension: independent 30 s windows per call
    pub fn wdr_state_set_lanes(st: *mut wdr_state, n_lanes: c_int) -> c_int;
    pub fn wdr_full_n_segments_from_state(st: *mut wdr_state) -> c_int;                           // :397
    pub fn wdr_full_get_segment_chunk_from_state(st: *mut wdr_state, i: c_int) -> c_int;
    pub fn wdr_full_get_segment_t0_from_state(st: *mut wdr_state, i: c_int) -> i64;               // start_timestamp(), :400
    pub fn wdr_full_get_segment_t1_from_state(st: *mut wdr_state, i: c_int) -> i64;               // end_timestamp()
    pub fn wdr_full_get_segment_text_from_state(st: *mut wdr_state, i: c_int) -> *const c_char;   // to_str()
    pub fn wdr_full_n_tokens_from_state(st: *mut wdr_state, i: c_int) -> c_int;                   // n_tokens(), :252
    pub fn wdr_full_get_token_text_from_state(ctx: *mut wdr_context, st: *mut wdr_state, i: c_int, j: c_int) -> *const c_char; // to_str_lossy(), :257
    pub fn wdr_full_get_token_data_from_state(st: *mut wdr_state, i: c_int, j: c_int) -> wdr_token_data; // token_data(), :272
    pub fn wdr_full_lang_id_from_state(st: *mut wdr_state) -> c_int;                              // :393
    pub fn wdr_lang_str(id: c_int) -> *const c_char;                                              // whisper_rs::get_lang_str, :394
    pub fn wdr_convert_integer_to_float_audio(pcm: *const i16, n: c_int, out: *mut c_float) -> c_int; // src/vad.rs:12
    // src/audio.rs:17-19: instead of bailing on a non-16 kHz / multi-channel WAV, resample on the device
    pub fn wdr_resample_n_out(n_frames: i64, sample_rate: c_int) -> i64;
    pub fn wdr_resample_i16(pcm: *const i16, n_frames: i64, channels: c_int, sample_rate: c_int, out_i16: *mut i16, out_f32: *mut c_float,
                            out_cap: i64, n_out: *mut i64) -> c_int;
    // whisper_tokenize (WhisperContext::tokenize): what whisper_full applies to `initial_prompt` (src/transcribe.rs:74-76, 383-386)
    pub fn wdr_tokenize(ctx: *mut wdr_context, text: *const c_char, tokens: *mut i32, n_max_tokens: c_int) -> c_int;
    // temperature ladder: the temperature whose result stands for chunk i of the last call (-1.0 if out of range)
    pub fn wdr_full_get_chunk_temperature_from_state(st: *mut wdr_state, i_chunk: c_int) -> c_float;
    // ---- Silero VAD (src/vad.rs:15-43) ----------------------------------------------------------------------------
    pub fn wdr_vad_default_context_params() -> wdr_vad_context_params;
    pub fn wdr_vad_default_params() -> wdr_vad_params;
    pub fn wdr_vad_init_from_file_with_params(path: *const c_char, p: wdr_vad_context_params) -> *mut wdr_vad; // WhisperVadContext::new
    pub fn wdr_vad_free(v: *mut wdr_vad);
    pub fn wdr_vad_segments_from_samples(v: *mut wdr_vad, p: wdr_vad_params, pcm: *const c_float, n: c_int) -> *mut wdr_vad_segments; // :31
    pub fn wdr_vad_segments_n(s: *mut wdr_vad_segments) -> c_int;
    pub fn wdr_vad_segments_get_segment_t0(s: *mut wdr_vad_segments, i: c_int) -> c_float;       // centiseconds, :40-43
    pub fn wdr_vad_segments_get_segment_t1(s: *mut wdr_vad_segments, i: c_int) -> c_float;
    pub fn wdr_vad_free_segments(s: *mut wdr_vad_segments);
    // ---- pyannote-rs surface --------------------------------------------------------------------------------------
    pub fn wdr_seg_init(path: *const c_char, seed: u64, device: c_int) -> *mut wdr_seg;           // the session get_segments opens
    pub fn wdr_seg_free(m: *mut wdr_seg);
    pub fn wdr_seg_get_segments(m: *mut wdr_seg, pcm: *const i16, n: i64) -> *mut wdr_seg_result; // pyannote_rs::get_segments, src/engine.rs:117-122
    pub fn wdr_seg_result_n(r: *mut wdr_seg_result) -> c_int;
    pub fn wdr_seg_result_start(r: *mut wdr_seg_result, i: c_int) -> c_double;
    pub fn wdr_seg_result_end(r: *mut wdr_seg_result, i: c_int) -> c_double;
    pub fn wdr_seg_result_samples(r: *mut wdr_seg_result, i: c_int, count: *mut i64) -> *const i16;
    pub fn wdr_seg_result_free(r: *mut wdr_seg_result);
    pub fn wdr_emb_init(path: *const c_char, seed: u64, device: c_int) -> *mut wdr_emb;           // EmbeddingExtractor::new, src/transcribe.rs:343
    pub fn wdr_emb_free(m: *mut wdr_emb);
    pub fn wdr_emb_dim(m: *mut wdr_emb) -> c_int;
    pub fn wdr_emb_compute_i16(m: *mut wdr_emb, pcm: *const i16, n: i64, out: *mut c_float) -> c_int; // compute(&samples), :466
    pub fn wdr_emb_compute_batch_i16(m: *mut wdr_emb, pcm: *const i16, seg_offset: *const i64, n_segments: c_int, out: *mut c_float,
                                     status: *mut i32) -> c_int;                                  // all segments of a recording in one launch sequence
    pub fn wdr_spk_init(max_speakers: size_t) -> *mut wdr_spk;                                    // EmbeddingManager::new, :342
    pub fn wdr_spk_free(m: *mut wdr_spk);
    pub fn wdr_spk_count(m: *mut wdr_spk) -> c_int;                                               // get_all_speakers().len(), :482
    pub fn wdr_spk_search(m: *mut wdr_spk, emb: *const c_float, dim: c_int, threshold: c_float) -> c_int; // search_speaker, :489 (0 = None)
    pub fn wdr_spk_best_match(m: *mut wdr_spk, emb: *const c_float, dim: c_int) -> c_int;         // get_best_speaker_match, :484
    pub fn wdr_cosine_matrix(emb: *const c_float, n: c_int, d: c_int, s: *mut c_float) -> c_int;
    pub fn wdr_cluster_leader(s: *const c_float, n: c_int, threshold: c_float, max_speakers: size_t, labels: *mut i32) -> c_int;
    pub fn wdr_cluster_agglomerative(s: *const c_float, n: c_int, threshold: c_float, labels: *mut i32) -> c_int;
    // the crate's per-segment speaker policy (src/transcribe.rs:480-492) over n embeddings in order; labels 0 = None -> "?"
    pub fn wdr_spk_assign_batch(m: *mut wdr_spk, emb: *const c_float, n: c_int, dim: c_int, threshold: c_float, labels: *mut i32) -> c_int;
    // ---- multi-GPU: the path's one exchange (the reference is single-GPU: `gpu_device`, src/engine.rs:14) ------------------------
    // one wdr_dist per GPU; n processes: rank 0 draws the id, every rank calls wdr_dist_init; one process: wdr_dist_init_all
    pub fn wdr_dist_available() -> c_int;
    pub fn wdr_dist_get_unique_id(id: *mut u8) -> c_int;
    pub fn wdr_dist_init(id: *const u8, n_ranks: c_int, rank: c_int, device: c_int) -> *mut wdr_dist;
    pub fn wdr_dist_init_all(n_gpus: c_int, devices: *const c_int, out: *mut *mut wdr_dist) -> c_int;
    pub fn wdr_dist_free(d: *mut wdr_dist);
    pub fn wdr_dist_size(d: *mut wdr_dist) -> c_int;
    pub fn wdr_dist_rank(d: *mut wdr_dist) -> c_int;
    pub fn wdr_allgather_embeddings(d: *mut wdr_dist, emb: *const c_float, n_local: c_int, dim: c_int, normalize: c_int, out: *mut c_float,
                                    out_cap_rows: i64, counts_out: *mut i32) -> c_int;
    // ---- model files read without a device (src/engine.rs:90-91, src/model_manager.rs:305-315) -----------------------------------
    pub fn wdr_onnx_probe(path: *const c_char, kind: c_int, info: *mut i32) -> c_int;
    pub fn wdr_silero_probe(path: *const c_char, hparams: *mut i32, n_tensors: *mut i32) -> c_int;
}

const _: () = assert!(std::mem::size_of::<wdr_token_data>() == 56);
