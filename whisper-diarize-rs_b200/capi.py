"""ctypes binding of libwdr_b200.so — one Python function per C entry point of include/wdr.h."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
_SO = os.path.join(_CSRC, "libwdr_b200.so")
_lib = None

f32p, i16p, i32p, i64p = C.POINTER(C.c_float), C.POINTER(C.c_int16), C.POINTER(C.c_int32), C.POINTER(C.c_int64)

WDR_ERR_NO_DEVICE = -2
WDR_ERR_TOO_SHORT = -6


class ContextParams(C.Structure):
    """== wdr_context_params (whisper_context_params + arch_name/seed)."""
    _fields_ = [("use_gpu", C.c_int), ("gpu_device", C.c_int), ("flash_attn", C.c_int), ("dtw_token_timestamps", C.c_int),
                ("dtw_aheads_preset", C.c_int), ("dtw_mem_size", C.c_size_t), ("arch_name", C.c_char_p), ("seed", C.c_uint64)]


class ModelDims(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("n_audio_state", "n_audio_head", "n_audio_layer", "n_text_layer", "n_mels", "n_vocab",
                                       "n_audio_ctx", "n_text_ctx", "is_multilingual")] + [("weight_bytes", C.c_int64)]


class TokenData(C.Structure):
    """== wdr_token_data == whisper_token_data (reference src/transcribe.rs:272-282)."""
    _fields_ = [("id", C.c_int32), ("tid", C.c_int32), ("p", C.c_float), ("plog", C.c_float), ("pt", C.c_float), ("ptsum", C.c_float),
                ("t0", C.c_int64), ("t1", C.c_int64), ("t_dtw", C.c_int64), ("vlen", C.c_float)]


PROGRESS_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p)
ABORT_CB = C.CFUNCTYPE(C.c_bool, C.c_void_p)


class FullParams(C.Structure):
    """== wdr_full_params (FullParams, reference src/transcribe.rs:20-87)."""
    _fields_ = [("strategy", C.c_int), ("n_threads", C.c_int), ("n_max_text_ctx", C.c_int), ("offset_ms", C.c_int), ("duration_ms", C.c_int),
                ("translate", C.c_int), ("no_context", C.c_int), ("no_timestamps", C.c_int), ("single_segment", C.c_int),
                ("print_special", C.c_int), ("print_progress", C.c_int), ("print_realtime", C.c_int), ("print_timestamps", C.c_int),
                ("token_timestamps", C.c_int), ("thold_pt", C.c_float), ("thold_ptsum", C.c_float), ("max_len", C.c_int),
                ("split_on_word", C.c_int), ("max_tokens", C.c_int), ("audio_ctx", C.c_int), ("initial_prompt", C.c_char_p),
                ("prompt_tokens", C.POINTER(C.c_int32)), ("prompt_n_tokens", C.c_int), ("language", C.c_char_p), ("detect_language", C.c_int),
                ("suppress_blank", C.c_int), ("suppress_nst", C.c_int), ("temperature", C.c_float), ("max_initial_ts", C.c_float),
                ("length_penalty", C.c_float), ("temperature_inc", C.c_float), ("entropy_thold", C.c_float), ("logprob_thold", C.c_float),
                ("no_speech_thold", C.c_float), ("greedy_best_of", C.c_int), ("beam_size", C.c_int), ("beam_patience", C.c_float),
                ("progress_callback", PROGRESS_CB), ("progress_callback_user_data", C.c_void_p), ("abort_callback", ABORT_CB),
                ("abort_callback_user_data", C.c_void_p)]


class VadContextParams(C.Structure):
    _fields_ = [("n_threads", C.c_int), ("use_gpu", C.c_int), ("gpu_device", C.c_int), ("seed", C.c_uint64)]


class VadParams(C.Structure):
    """== wdr_vad_params == whisper_vad_params (WhisperVadParams, reference src/vad.rs:21-27)."""
    _fields_ = [("threshold", C.c_float), ("min_speech_duration_ms", C.c_int), ("min_silence_duration_ms", C.c_int),
                ("max_speech_duration_s", C.c_float), ("speech_pad_ms", C.c_int), ("samples_overlap", C.c_float)]


class WdrError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"wdr error {code}: {msg}")
        self.code = code


def lib_path():
    return _SO


def build(force=False):
    """Compile csrc/*.cu for sm_100a into csrc/libwdr_b200.so (nvcc cross-compiles without a GPU)."""
    if force:
        subprocess.check_call(["make", "-C", _CSRC, "clean"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", _CSRC, "-j8", "-s"])
    return _SO


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        raise WdrError(-100, f"{_SO} is missing: run __graft_entry__.build() (there is no CPU fallback)")
    L = C.CDLL(_SO)
    L.wdr_version.restype = C.c_char_p
    L.wdr_last_error.restype = C.c_char_p
    L.wdr_launch_count.restype = C.c_uint64
    L.wdr_mel_init.restype = C.c_void_p
    L.wdr_mel_init.argtypes = [f32p, C.c_int, C.c_int]
    L.wdr_mel_free.argtypes = [C.c_void_p]
    L.wdr_mel_n_len.argtypes = [C.c_int]
    L.wdr_log_mel_f32.argtypes = [C.c_void_p, f32p, C.c_int, C.c_int, f32p]
    L.wdr_log_mel_i16.argtypes = [C.c_void_p, i16p, C.c_int, C.c_int, f32p]
    L.wdr_log_mel_batch_f32_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.wdr_log_mel_batch_i16_dev.argtypes = L.wdr_log_mel_batch_f32_dev.argtypes
    L.wdr_log_mel_batch_i16.argtypes = [C.c_void_p, i16p, C.c_int64, i32p, C.c_int, C.c_int, f32p]
    L.wdr_convert_integer_to_float_audio.argtypes = [i16p, C.c_int, f32p]
    L.wdr_resample_n_out.argtypes = [C.c_int64, C.c_int]
    L.wdr_resample_n_out.restype = C.c_int64
    L.wdr_resample_i16.argtypes = [i16p, C.c_int64, C.c_int, C.c_int, i16p, f32p, C.c_int64, C.POINTER(C.c_int64)]
    L.wdr_median_filter.argtypes = [f32p, C.c_int, C.c_int, C.c_int, C.c_int, f32p]
    L.wdr_dtw_cost.argtypes = [f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, f32p]
    L.wdr_dtw.argtypes = [f32p, C.c_int, C.c_int, i32p, i32p, C.POINTER(C.c_int), f32p, i32p]
    L.wdr_dtw_batch_dev.argtypes = [C.c_void_p, i64p, i32p, i32p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.wdr_fbank_frames.argtypes = [C.c_int]
    L.wdr_kaldi_fbank_i16.argtypes = [i16p, C.c_int, C.c_int, C.c_int, f32p]
    L.wdr_kaldi_fbank_batch_i16_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.wdr_signal_energy.argtypes = [f32p, C.c_int, C.c_int, f32p]
    L.wdr_gemm_bf16_dev.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_int,
                                    C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64,
                                    C.c_int, C.c_int64, C.c_void_p]
    L.wdr_context_default_params.restype = ContextParams
    L.wdr_init_from_file_with_params.restype = C.c_void_p
    L.wdr_init_from_file_with_params.argtypes = [C.c_char_p, ContextParams]
    L.wdr_free.argtypes = [C.c_void_p]
    L.wdr_model_info.argtypes = [C.c_void_p, C.POINTER(ModelDims)]
    L.wdr_init_state.restype = C.c_void_p
    L.wdr_ggml_probe.argtypes = [C.c_char_p, i32p, i32p, i32p]
    L.wdr_onnx_probe.argtypes = [C.c_char_p, C.c_int, i32p]
    L.wdr_onnx_read_param.argtypes = [C.c_char_p, C.c_int, C.c_char_p, f32p, C.c_int64]
    L.wdr_onnx_read_param.restype = C.c_int64
    L.wdr_silero_probe.argtypes = [C.c_char_p, i32p, i32p]
    L.wdr_init_state.argtypes = [C.c_void_p]
    L.wdr_free_state.argtypes = [C.c_void_p]
    L.wdr_mel_filters.argtypes = [C.c_int, f32p]
    L.wdr_encode.argtypes = [C.c_void_p, C.c_void_p, f32p, C.c_int, C.c_int, f32p]
    L.wdr_encode_chunks_i16_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.wdr_encode_chunks_i16.argtypes = [C.c_void_p, C.c_void_p, i16p, C.c_int64, i32p, C.c_int, f32p]
    L.wdr_state_hidden_digest.argtypes = [C.c_void_p, f32p, C.c_int]
    L.wdr_profile_enable.argtypes = [C.c_void_p, C.c_int]
    L.wdr_profile_collect.argtypes = [C.c_void_p, C.POINTER(C.c_double), i32p, C.c_int]
    L.wdr_state_set_lanes.argtypes = [C.c_void_p, C.c_int]
    L.wdr_encoder_attention_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.wdr_full_default_params.restype = FullParams
    L.wdr_full_default_params.argtypes = [C.c_int]
    L.wdr_full_with_state.argtypes = [C.c_void_p, C.c_void_p, FullParams, f32p, C.c_int]
    L.wdr_full_with_state_i16.argtypes = [C.c_void_p, C.c_void_p, FullParams, i16p, C.c_int]
    L.wdr_full_batch_i16.argtypes = [C.c_void_p, C.c_void_p, FullParams, C.c_void_p, C.c_int64, i32p, C.c_int]
    L.wdr_full_batch_i16_dev.argtypes = [C.c_void_p, C.c_void_p, FullParams, C.c_void_p, C.c_int64, i32p, C.c_int]
    L.wdr_full_n_segments_from_state.argtypes = [C.c_void_p]
    for fn in ("chunk", "t0", "t1", "text", "no_speech_prob"):
        getattr(L, f"wdr_full_get_segment_{fn}_from_state").argtypes = [C.c_void_p, C.c_int]
    L.wdr_full_get_segment_t0_from_state.restype = C.c_int64
    L.wdr_full_get_segment_t1_from_state.restype = C.c_int64
    L.wdr_full_get_segment_text_from_state.restype = C.c_char_p
    L.wdr_full_get_segment_no_speech_prob_from_state.restype = C.c_float
    L.wdr_full_n_tokens_from_state.argtypes = [C.c_void_p, C.c_int]
    L.wdr_full_get_token_id_from_state.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.wdr_full_get_token_text_from_state.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    L.wdr_full_get_token_text_from_state.restype = C.c_char_p
    L.wdr_full_get_token_data_from_state.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.wdr_full_get_token_data_from_state.restype = TokenData
    L.wdr_full_lang_id_from_state.argtypes = [C.c_void_p]
    L.wdr_full_get_chunk_lang_id_from_state.argtypes = [C.c_void_p, C.c_int]
    L.wdr_lang_str.argtypes = [C.c_int]
    L.wdr_lang_str.restype = C.c_char_p
    L.wdr_lang_id.argtypes = [C.c_char_p]
    L.wdr_token_to_str.argtypes = [C.c_void_p, C.c_int32]
    L.wdr_token_to_str.restype = C.c_char_p
    L.wdr_full_get_chunk_info_from_state.argtypes = [C.c_void_p, C.c_int, i32p, f32p]
    L.wdr_sample_discrete.argtypes = [f32p, C.c_int, C.c_uint32, C.c_int, i32p]
    L.wdr_tokenize.argtypes = [C.c_void_p, C.c_char_p, i32p, C.c_int]
    L.wdr_tokenize_with_vocab.argtypes = [C.POINTER(C.c_char_p), C.c_int, C.c_char_p, i32p, C.c_int]
    L.wdr_full_get_chunk_temperature_from_state.argtypes = [C.c_void_p, C.c_int]
    L.wdr_full_get_chunk_temperature_from_state.restype = C.c_float
    L.wdr_decode_teacher_forced.argtypes = [C.c_void_p, C.c_void_p, f32p, C.c_int, i32p, C.c_int, f32p, f32p]
    L.wdr_vad_default_context_params.restype = VadContextParams
    L.wdr_vad_default_params.restype = VadParams
    L.wdr_vad_init_from_file_with_params.restype = C.c_void_p
    L.wdr_vad_init_from_file_with_params.argtypes = [C.c_char_p, VadContextParams]
    L.wdr_vad_free.argtypes = [C.c_void_p]
    L.wdr_vad_detect_speech.argtypes = [C.c_void_p, f32p, C.c_int]
    L.wdr_vad_n_probs.argtypes = [C.c_void_p]
    L.wdr_vad_probs.argtypes = [C.c_void_p]
    L.wdr_vad_probs.restype = f32p
    L.wdr_vad_detect_speech_batch_i16.argtypes = [C.c_void_p, i16p, i64p, i32p, C.c_int, f32p, i64p]
    L.wdr_vad_segments_from_probs.restype = C.c_void_p
    L.wdr_vad_segments_from_probs.argtypes = [C.c_void_p, VadParams]
    L.wdr_vad_segments_from_probs_array.restype = C.c_void_p
    L.wdr_vad_segments_from_probs_array.argtypes = [f32p, C.c_int, VadParams]
    L.wdr_vad_segments_from_samples.restype = C.c_void_p
    L.wdr_vad_segments_from_samples.argtypes = [C.c_void_p, VadParams, f32p, C.c_int]
    L.wdr_vad_segments_n.argtypes = [C.c_void_p]
    L.wdr_vad_segments_get_segment_t0.argtypes = [C.c_void_p, C.c_int]
    L.wdr_vad_segments_get_segment_t0.restype = C.c_float
    L.wdr_vad_segments_get_segment_t1.argtypes = [C.c_void_p, C.c_int]
    L.wdr_vad_segments_get_segment_t1.restype = C.c_float
    L.wdr_vad_free_segments.argtypes = [C.c_void_p]
    L.wdr_seg_init.restype = C.c_void_p
    L.wdr_seg_init.argtypes = [C.c_char_p, C.c_uint64, C.c_int]
    L.wdr_seg_free.argtypes = [C.c_void_p]
    L.wdr_full_get_cross_attn_stats.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    u8p = C.POINTER(C.c_uint8)
    L.wdr_dist_get_unique_id.argtypes = [u8p]
    L.wdr_dist_init.restype = C.c_void_p
    L.wdr_dist_init.argtypes = [u8p, C.c_int, C.c_int, C.c_int]
    L.wdr_dist_free.argtypes = [C.c_void_p]
    L.wdr_dist_size.argtypes = [C.c_void_p]
    L.wdr_dist_rank.argtypes = [C.c_void_p]
    L.wdr_allgather_embeddings.argtypes = [C.c_void_p, f32p, C.c_int, C.c_int, C.c_int, f32p, C.c_int64, i32p]
    L.wdr_allgather_embeddings_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64, i32p, C.c_void_p]
    L.wdr_spk_assign_batch.argtypes = [C.c_void_p, f32p, C.c_int, C.c_int, C.c_float, i32p]
    L.wdr_gemm_last_tile_n.restype = C.c_int
    L.wdr_gemm_last_tile_n.argtypes = []
    L.wdr_seg_n_windows.argtypes = [C.c_int64]
    L.wdr_seg_scores_i16.argtypes = [C.c_void_p, i16p, C.c_int64, f32p]
    L.wdr_seg_get_segments.restype = C.c_void_p
    L.wdr_seg_get_segments.argtypes = [C.c_void_p, i16p, C.c_int64]
    L.wdr_seg_segments_from_scores.restype = C.c_void_p
    L.wdr_seg_segments_from_scores.argtypes = [f32p, C.c_int, C.c_int64]
    L.wdr_seg_result_n.argtypes = [C.c_void_p]
    L.wdr_seg_result_start.argtypes = [C.c_void_p, C.c_int]
    L.wdr_seg_result_start.restype = C.c_double
    L.wdr_seg_result_end.argtypes = [C.c_void_p, C.c_int]
    L.wdr_seg_result_end.restype = C.c_double
    L.wdr_seg_result_sample_range.argtypes = [C.c_void_p, C.c_int, i64p]
    L.wdr_seg_result_sample_range.restype = C.c_int64
    L.wdr_seg_result_samples.argtypes = [C.c_void_p, C.c_int, i64p]
    L.wdr_seg_result_samples.restype = i16p
    L.wdr_seg_result_free.argtypes = [C.c_void_p]
    L.wdr_emb_init.restype = C.c_void_p
    L.wdr_emb_init.argtypes = [C.c_char_p, C.c_uint64, C.c_int]
    L.wdr_emb_free.argtypes = [C.c_void_p]
    L.wdr_emb_dim.argtypes = [C.c_void_p]
    L.wdr_emb_compute_i16.argtypes = [C.c_void_p, i16p, C.c_int64, f32p]
    L.wdr_emb_compute_batch_i16.argtypes = [C.c_void_p, i16p, i64p, C.c_int, f32p, i32p]
    L.wdr_emb_compute_batch_i16_dev.argtypes = [C.c_void_p, C.c_void_p, i64p, C.c_int, C.c_void_p, i32p, C.c_void_p]
    L.wdr_emb_last_flops.argtypes = [C.c_void_p]
    L.wdr_emb_last_flops.restype = C.c_double
    L.wdr_emb_profile.argtypes = [C.c_void_p, C.c_int]
    L.wdr_emb_last_kernel_ms.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.wdr_full_get_phase_ms.argtypes = [C.c_void_p, C.POINTER(C.c_double), i32p]
    L.wdr_spk_init.restype = C.c_void_p
    L.wdr_spk_init.argtypes = [C.c_size_t]
    L.wdr_spk_free.argtypes = [C.c_void_p]
    L.wdr_spk_count.argtypes = [C.c_void_p]
    L.wdr_spk_search.argtypes = [C.c_void_p, f32p, C.c_int, C.c_float]
    L.wdr_spk_best_match.argtypes = [C.c_void_p, f32p, C.c_int]
    L.wdr_cosine_matrix.argtypes = [f32p, C.c_int, C.c_int, f32p]
    L.wdr_cluster_leader.argtypes = [f32p, C.c_int, C.c_float, C.c_size_t, i32p]
    L.wdr_cluster_agglomerative.argtypes = [f32p, C.c_int, C.c_float, i32p]
    _lib = L
    return L


def _check(rc):
    if rc < 0:
        raise WdrError(rc, load().wdr_last_error().decode("utf-8", "replace"))
    return rc


def _np(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def _p(a, ptype):
    return a.ctypes.data_as(ptype)


def version():
    return load().wdr_version().decode()


def device_count():
    return load().wdr_device_count()


def launch_count():
    return int(load().wdr_launch_count())


def mel_n_len(n):
    return load().wdr_mel_n_len(int(n))


class MelFrontend:
    """wdr_mel handle: whisper.cpp log_mel_spectrogram on the device (reference src/transcribe.rs:389)."""

    def __init__(self, filters, device=0):
        f = _np(filters, np.float32)
        assert f.ndim == 2 and f.shape[1] == 201
        self.n_mel = f.shape[0]
        self._h = load().wdr_mel_init(_p(f, f32p), self.n_mel, device)
        if not self._h:
            raise WdrError(WDR_ERR_NO_DEVICE if device_count() == 0 else -3, load().wdr_last_error().decode())

    def close(self):
        if self._h:
            load().wdr_mel_free(self._h)
            self._h = None

    __del__ = close

    def log_mel(self, pcm, normalize=True):
        """Whole buffer, whisper.cpp layout: returns mel[n_mel, n_len]."""
        pcm = np.asarray(pcm)
        n_len = mel_n_len(len(pcm))
        out = np.empty((self.n_mel, n_len), np.float32)
        if pcm.dtype == np.int16:
            x = _np(pcm, np.int16)
            _check(load().wdr_log_mel_i16(self._h, _p(x, i16p), len(x), int(normalize), _p(out, f32p)))
        else:
            x = _np(pcm, np.float32)
            _check(load().wdr_log_mel_f32(self._h, _p(x, f32p), len(x), int(normalize), _p(out, f32p)))
        return out

    def log_mel_batch(self, pcm_i16, n_valid=None, normalize=True):
        """pcm_i16[B, 480000] host array -> mel[B, n_mel, 3000] (H2D + kernel + D2H inside the call)."""
        x = _np(pcm_i16, np.int16)
        assert x.ndim == 2 and x.shape[1] == 480000
        out = np.empty((x.shape[0], self.n_mel, 3000), np.float32)
        nv = None if n_valid is None else _np(n_valid, np.int32)
        _check(load().wdr_log_mel_batch_i16(self._h, _p(x, i16p), x.shape[1], None if nv is None else _p(nv, i32p),
                                            x.shape[0], int(normalize), _p(out, f32p)))
        return out

    def log_mel_batch_dev(self, pcm_ptr, is_i16, chunk_stride, n_chunks, out_ptr, n_valid_ptr=None, out_max_ptr=None,
                          normalize=True, stream=0):
        """Device pointers (ints), asynchronous on `stream` (a cudaStream_t value)."""
        fn = load().wdr_log_mel_batch_i16_dev if is_i16 else load().wdr_log_mel_batch_f32_dev
        _check(fn(self._h, pcm_ptr, chunk_stride, n_valid_ptr, n_chunks, int(normalize), out_ptr, out_max_ptr, stream))


def log_mel(pcm, filters, normalize=True, device=0):
    m = MelFrontend(filters, device)
    try:
        return m.log_mel(pcm, normalize)
    finally:
        m.close()


def convert_integer_to_float_audio(pcm_i16):
    x = _np(pcm_i16, np.int16)
    out = np.empty(len(x), np.float32)
    _check(load().wdr_convert_integer_to_float_audio(_p(x, i16p), len(x), _p(out, f32p)))
    return out


def resample_to_16k(pcm_i16, sample_rate, channels=1, want_f32=False):
    """interleaved int16 at `sample_rate` -> 16 kHz mono int16 (and the unrounded f32 / 32768 if want_f32): wdr_resample_i16."""
    x = _np(pcm_i16, np.int16).reshape(-1)
    n_frames = len(x) // channels
    n = int(load().wdr_resample_n_out(n_frames, int(sample_rate)))
    if n < 0:
        raise ValueError("bad sample rate")
    o16 = np.empty(n, np.int16)
    o32 = np.empty(n, np.float32) if want_f32 else None
    n_out = C.c_int64(0)
    _check(load().wdr_resample_i16(_p(x, i16p), n_frames, channels, int(sample_rate), _p(o16, i16p),
                                   _p(o32, f32p) if want_f32 else None, n, C.byref(n_out)))
    assert n_out.value == n
    return (o16, o32) if want_f32 else o16


def tokenize_with_vocab(token_strings, text, n_max=4096):
    """whisper_tokenize over an explicit vocabulary (list of str / bytes / None by id): wdr_tokenize_with_vocab."""
    arr = (C.c_char_p * len(token_strings))(*[None if t is None else (t if isinstance(t, bytes) else t.encode()) for t in token_strings])
    out = np.empty(n_max, np.int32)
    n = load().wdr_tokenize_with_vocab(arr, len(token_strings), text.encode() if isinstance(text, str) else text, _p(out, i32p), n_max)
    if n < 0:
        raise ValueError(f"{-n} tokens do not fit {n_max}")
    return out[:n].copy()


def sample_discrete(logprobs, seed, n_draws):
    """n_draws consecutive whisper_sample_token(best=false) draws (host-side sampler of the temperature ladder)."""
    lp = _np(logprobs, np.float32)
    ids = np.empty(n_draws, np.int32)
    _check(load().wdr_sample_discrete(_p(lp, f32p), len(lp), int(seed), int(n_draws), _p(ids, i32p)))
    return ids


def median_filter(w, width=7):
    w3 = _np(w, np.float32)
    H, N, M = w3.shape
    out = np.empty_like(w3)
    _check(load().wdr_median_filter(_p(w3, f32p), H, N, M, width, _p(out, f32p)))
    return out


def dtw_cost(w, sot_len, width=7):
    w3 = _np(w, np.float32)
    H, T, A = w3.shape
    out = np.empty((T - sot_len - 1, A), np.float32)
    _check(load().wdr_dtw_cost(_p(w3, f32p), H, T, A, sot_len, width, _p(out, f32p)))
    return out


def dtw(x, want_matrices=False):
    x2 = _np(x, np.float32)
    N, M = x2.shape
    ti = np.empty(N + M + 2, np.int32)
    tj = np.empty(N + M + 2, np.int32)
    n = C.c_int(0)
    cost = trace = None
    cp, tp = f32p(), i32p()
    if want_matrices:
        cost = np.empty((N + 1, M + 1), np.float32)
        trace = np.empty((N + 1, M + 1), np.int32)
        cp, tp = _p(cost, f32p), _p(trace, i32p)
    _check(load().wdr_dtw(_p(x2, f32p), N, M, _p(ti, i32p), _p(tj, i32p), C.byref(n), cp, tp))
    if want_matrices:
        return ti[: n.value].copy(), tj[: n.value].copy(), cost, trace
    return ti[: n.value].copy(), tj[: n.value].copy()


def dtw_batch_dev(x_ptr, x_offset, N, M, text_ptr, time_ptr, len_ptr, max_path, stream=0):
    xo, n_, m_ = _np(x_offset, np.int64), _np(N, np.int32), _np(M, np.int32)
    _check(load().wdr_dtw_batch_dev(x_ptr, _p(xo, i64p), _p(n_, i32p), _p(m_, i32p), len(n_), text_ptr, time_ptr, len_ptr,
                                    max_path, stream))


def fbank_frames(n):
    return load().wdr_fbank_frames(int(n))


def kaldi_fbank(pcm_i16, n_bins=80, subtract_mean=True):
    x = _np(pcm_i16, np.int16)
    T = fbank_frames(len(x))
    out = np.empty((max(T, 0), n_bins), np.float32)
    rc = _check(load().wdr_kaldi_fbank_i16(_p(x, i16p), len(x), n_bins, int(subtract_mean), _p(out, f32p)))
    assert rc == T
    return out


def signal_energy(pcm_f32, hw=32):
    x = _np(pcm_f32, np.float32)
    out = np.empty_like(x)
    _check(load().wdr_signal_energy(_p(x, f32p), len(x), hw, _p(out, f32p)))
    return out


def gemm_bf16_dev(A_ptr, lda, rows_per_batch, n_batch, a_batch_stride, W_ptr, ldw, N, K, out_ptr, ldc, epilogue=0, bias_ptr=None,
                  extra_ptr=None, out_t_ptr=None, ldt=0, n_split=0, kb_per_tap=0, a_cols=0, t_batch_stride=0, stream=0):
    """tcgen05 GEMM on device pointers (ints); see wdr_gemm_bf16_dev in include/wdr.h."""
    _check(load().wdr_gemm_bf16_dev(A_ptr, lda, rows_per_batch, n_batch, a_batch_stride, W_ptr, ldw, N, K, kb_per_tap, a_cols,
                                    bias_ptr, epilogue, out_ptr, ldc, extra_ptr, out_t_ptr, ldt, n_split, t_batch_stride, stream))


def mel_filters(n_mel):
    out = np.empty((n_mel, 201), np.float32)
    _check(load().wdr_mel_filters(n_mel, _p(out, f32p)))
    return out


DTW_PRESETS = {"tiny.en": 0, "tiny": 1, "base.en": 2, "base": 3, "small.en": 4, "small": 5, "medium.en": 6, "medium": 7,
               "large-v3": 8, "large-v3-turbo": 9}


class Context:
    """wdr_context: WhisperContext::new_with_params (reference src/transcribe.rs:89-166)."""

    def __init__(self, arch_name, seed=1234, gpu_device=0, enable_dtw=False, flash_attn=False, dtw_mem_size=0, model_path=None):
        """arch_name: the model name the crate maps to a DTW preset (src/transcribe.rs:117-129).  model_path: a ggml-<model>.bin
        checkpoint (None = seeded synthetic weights of `arch_name`)."""
        L = load()
        p = L.wdr_context_default_params()
        p.gpu_device = gpu_device
        p.arch_name = arch_name.encode()
        p.seed = seed
        if enable_dtw:  # create_context: DTW on => flash attention off, preset by model name, unknown names -> small
            p.flash_attn = 0
            p.dtw_token_timestamps = 1
            p.dtw_aheads_preset = DTW_PRESETS.get(arch_name, DTW_PRESETS["small"])
            p.dtw_mem_size = dtw_mem_size
        else:
            p.flash_attn = int(flash_attn)
        self._h = L.wdr_init_from_file_with_params(model_path.encode() if model_path else None, p)
        if not self._h:
            raise WdrError(WDR_ERR_NO_DEVICE if device_count() == 0 else -3, L.wdr_last_error().decode())
        self.arch_name = arch_name
        md = ModelDims()
        _check(L.wdr_model_info(self._h, C.byref(md)))
        self.dims = md

    def close(self):
        if self._h:
            load().wdr_free(self._h)
            self._h = None

    __del__ = close

    def create_state(self):
        return State(self)

    def tokenize(self, text, n_max=4096):
        """whisper_tokenize with this context's vocabulary (what whisper_full applies to `initial_prompt`)."""
        out = np.empty(n_max, np.int32)
        n = load().wdr_tokenize(self._h, text.encode() if isinstance(text, str) else text, _p(out, i32p), n_max)
        if n < 0:
            raise ValueError(f"{-n} tokens do not fit {n_max}")
        return out[:n].copy()


class State:
    """wdr_state: ctx.create_state() (reference src/transcribe.rs:335)."""

    def __init__(self, ctx):
        self.ctx = ctx
        self._h = load().wdr_init_state(ctx._h)
        if not self._h:
            raise WdrError(-3, load().wdr_last_error().decode())

    def close(self):
        if self._h:
            load().wdr_free_state(self._h)
            self._h = None

    __del__ = close

    def encode(self, mel, mel_offset=0):
        """whisper_encode: normalised mel[n_mel, n_len] -> hidden[1500, d]."""
        m = _np(mel, np.float32)
        out = np.empty((1500, self.ctx.dims.n_audio_state), np.float32)
        _check(load().wdr_encode(self.ctx._h, self._h, _p(m, f32p), m.shape[1], mel_offset, _p(out, f32p)))
        return out

    def encode_chunks(self, pcm_i16, n_valid=None):
        """pcm_i16[B, 480000] (host) -> hidden[B, 1500, d]; H2D/D2H inside the call."""
        x = _np(pcm_i16, np.int16)
        assert x.ndim == 2 and x.shape[1] == 480000
        out = np.empty((x.shape[0], 1500, self.ctx.dims.n_audio_state), np.float32)
        nv = None if n_valid is None else _np(n_valid, np.int32)
        _check(load().wdr_encode_chunks_i16(self.ctx._h, self._h, _p(x, i16p), x.shape[1], None if nv is None else _p(nv, i32p),
                                            x.shape[0], _p(out, f32p)))
        return out

    def encode_chunks_resident(self, pcm_host_ptr, n_chunks, chunk_stride=480000):
        """Host PCM pointer (int; pinned memory recommended) -> encoder output kept in the state on the device."""
        _check(load().wdr_encode_chunks_i16(self.ctx._h, self._h, C.cast(pcm_host_ptr, i16p), chunk_stride, None, n_chunks, None))

    def hidden_digest(self, n):
        out = np.empty(n, np.float32)
        _check(load().wdr_state_hidden_digest(self._h, _p(out, f32p), n))
        return out

    def profile_enable(self, on=True):
        _check(load().wdr_profile_enable(self._h, int(on)))

    def set_lanes(self, n):
        """Lanes of full_batch (0 = library default); output is identical for any lane count."""
        _check(load().wdr_state_set_lanes(self._h, int(n)))

    def profile_collect(self):
        ms = (C.c_double * 16)()
        ln = (C.c_int32 * 16)()
        _check(load().wdr_profile_collect(self._h, ms, ln, 16))
        names = ["mel", "mel_aux", "gemm", "attention", "layernorm", "decoder", "dtw", "other", "dec_cross", "dec_gemm", "dec_cross_batched"]
        return {n: {"ms": ms[i], "records": ln[i]} for i, n in enumerate(names)}

    # ---- full transcription (state.full, reference src/transcribe.rs:389; accessors :393-412, :252-282) ----
    def full_params(self, strategy=0, temperature_inc=0.0, **kw):
        """FullParams as setup_params builds them (src/transcribe.rs:20-87): suppress_blank, token_timestamps, single_segment.
        wdr_full_default_params returns whisper.cpp's temperature_inc = 0.2 (what the crate runs with); THIS helper defaults to 0 (no
        fallback ladder) because a random-init model fails the log-probability test on every window — pass temperature_inc=0.2 for
        the crate's default behaviour."""
        p = load().wdr_full_default_params(strategy)
        p.temperature_inc = temperature_inc
        p.print_special = 0
        p.print_progress = 1
        p.print_realtime = 0
        p.print_timestamps = 0
        p.suppress_blank = 1
        p.token_timestamps = 1
        p.single_segment = 1
        for k, v in kw.items():
            setattr(p, k, v.encode() if isinstance(v, str) else v)
        return p

    def full(self, pcm, params=None):
        """state.full on ONE buffer (float32 in [-1,1) or int16); longer than 30 s = whisper_full's sequential seek loop."""
        p = params if params is not None else self.full_params()
        pcm = np.asarray(pcm)
        if pcm.dtype == np.int16:
            x = _np(pcm, np.int16)
            _check(load().wdr_full_with_state_i16(self.ctx._h, self._h, p, _p(x, i16p), len(x)))
        else:
            x = _np(pcm, np.float32)
            _check(load().wdr_full_with_state(self.ctx._h, self._h, p, _p(x, f32p), len(x)))
        return self.segments()

    def full_batch(self, pcm_i16, n_valid=None, params=None):
        """Sharded mode: pcm_i16[B, stride] host array (or a host pointer + (n_chunks, stride)) -> segments of all chunks."""
        p = params if params is not None else self.full_params()
        x = _np(pcm_i16, np.int16)
        assert x.ndim == 2
        nv = None if n_valid is None else _np(n_valid, np.int32)
        _check(load().wdr_full_batch_i16(self.ctx._h, self._h, p, x.ctypes.data, x.shape[1], None if nv is None else _p(nv, i32p), x.shape[0]))
        return self.segments()

    def full_batch_ptr(self, pcm_host_ptr, n_chunks, chunk_stride=480000, params=None):
        p = params if params is not None else self.full_params()
        _check(load().wdr_full_batch_i16(self.ctx._h, self._h, p, pcm_host_ptr, chunk_stride, None, n_chunks))
        return load().wdr_full_n_segments_from_state(self._h)

    def full_batch_dev(self, pcm_dev_ptr, n_chunks, chunk_stride=480000, params=None, n_valid=None):
        """PCM already in HBM (device pointer int); results gathered into the state.  Returns the number of segments."""
        p = params if params is not None else self.full_params()
        nv = None if n_valid is None else _np(n_valid, np.int32)
        _check(load().wdr_full_batch_i16_dev(self.ctx._h, self._h, p, pcm_dev_ptr, chunk_stride, None if nv is None else _p(nv, i32p), n_chunks))
        return load().wdr_full_n_segments_from_state(self._h)

    def n_segments(self):
        return load().wdr_full_n_segments_from_state(self._h)

    def segments(self):
        L = load()
        out = []
        for i in range(L.wdr_full_n_segments_from_state(self._h)):
            n = L.wdr_full_n_tokens_from_state(self._h, i)
            toks = [L.wdr_full_get_token_data_from_state(self._h, i, j) for j in range(n)]
            texts = [L.wdr_full_get_token_text_from_state(self.ctx._h, self._h, i, j).decode("utf-8", "replace") for j in range(n)]
            out.append(dict(chunk=L.wdr_full_get_segment_chunk_from_state(self._h, i), t0=L.wdr_full_get_segment_t0_from_state(self._h, i),
                            t1=L.wdr_full_get_segment_t1_from_state(self._h, i),
                            text=L.wdr_full_get_segment_text_from_state(self._h, i).decode("utf-8", "replace"),
                            no_speech_prob=L.wdr_full_get_segment_no_speech_prob_from_state(self._h, i), tokens=toks, token_text=texts))
        return out

    def chunk_info(self, i):
        info = np.zeros(8, np.int32)
        nsp = C.c_float(0)
        _check(load().wdr_full_get_chunk_info_from_state(self._h, i, _p(info, i32p), C.byref(nsp)))
        keys = ["seek_delta", "failed", "completed", "n_sampled", "has_ts", "result_len", "seek_end", "n_segments"]
        d = {k: int(v) for k, v in zip(keys, info)}
        d["no_speech_prob"] = float(nsp.value)
        d["temperature"] = float(load().wdr_full_get_chunk_temperature_from_state(self._h, i))
        return d

    def lang_id(self):
        return load().wdr_full_lang_id_from_state(self._h)

    def chunk_lang_id(self, i):
        """Language detected / used for buffer i of the last full call."""
        return load().wdr_full_get_chunk_lang_id_from_state(self._h, int(i))

    def cross_attn_stats(self):
        """(launches, live (launch, window) pairs) of dec_cross_attn_kernel in the last full call, counted on the device."""
        a, b = C.c_int64(0), C.c_int64(0)
        _check(load().wdr_full_get_cross_attn_stats(self._h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def phase_ms(self):
        """Device time of the phases of the last full call (summed over groups and lanes) + greedy iterations run."""
        ms = (C.c_double * 5)()
        n = C.c_int32(0)
        _check(load().wdr_full_get_phase_ms(self._h, ms, C.byref(n)))
        d = dict(zip(["encode", "cross_kv", "decode", "dtw_pass", "dtw"], [float(v) for v in ms]))
        d["decode_steps"] = int(n.value)
        return d

    def decode_teacher_forced(self, seq, enc=None, want_logits=True, want_aheads=False, n_aheads=0):
        """Stage-level: teacher-forced decoder pass. seq[B, n_seq]; enc[B,1500,d] host (None = keep the state's encoder output)."""
        sq = _np(seq, np.int32)
        B, n_seq = sq.shape
        nv = self.ctx.dims.n_vocab
        e = None if enc is None else _np(enc, np.float32)
        logits = np.empty((B, n_seq, nv), np.float32) if want_logits else None
        ah = np.empty((B, n_aheads, n_seq, 1500), np.float32) if want_aheads else None
        _check(load().wdr_decode_teacher_forced(self.ctx._h, self._h, None if e is None else _p(e, f32p), B, _p(sq, i32p), n_seq,
                                                None if logits is None else _p(logits, f32p), None if ah is None else _p(ah, f32p)))
        return logits, ah

    def encode_chunks_dev(self, pcm_ptr, chunk_stride, n_chunks, out_ptr, n_valid_ptr=None, stream=0):
        _check(load().wdr_encode_chunks_i16_dev(self.ctx._h, self._h, pcm_ptr, chunk_stride, n_valid_ptr, n_chunks, out_ptr, stream))


def encoder_attention_dev(qk_ptr, vt_ptr, ldt, n_chunks, T, n_head, d_model, out_ptr, stream=0):
    _check(load().wdr_encoder_attention_dev(qk_ptr, vt_ptr, ldt, n_chunks, T, n_head, d_model, out_ptr, stream))


def ggml_probe(path):
    """Header of a ggml checkpoint (no GPU needed): dict of the 11 hyper-parameters + n_tensors + n_tokens."""
    hp = np.zeros(11, np.int32)
    nt, nk = C.c_int32(0), C.c_int32(0)
    _check(load().wdr_ggml_probe(path.encode(), _p(hp, i32p), C.byref(nt), C.byref(nk)))
    keys = ["n_vocab", "n_audio_ctx", "n_audio_state", "n_audio_head", "n_audio_layer", "n_text_ctx", "n_text_state", "n_text_head",
            "n_text_layer", "n_mels", "ftype"]
    d = {k: int(v) for k, v in zip(keys, hp)}
    d.update(n_tensors=int(nt.value), n_tokens=int(nk.value))
    return d


ONNX_PYANNET, ONNX_RESNET34 = 0, 1


def onnx_probe(path, kind):
    """What the dependency-free ONNX reader finds in `path` and takes from it for network `kind` (no GPU needed)."""
    info = np.zeros(6, np.int32)
    _check(load().wdr_onnx_probe(path.encode(), kind, _p(info, i32p)))
    return dict(zip(["nodes", "tensors", "inputs", "outputs", "params", "emb_dim"], (int(v) for v in info)))


def onnx_read_param(path, kind, name):
    """One parameter as the loader extracts it (PyTorch name / layout; LSTM gates re-ordered; BatchNorm folded), flat fp32."""
    L = load()
    n = L.wdr_onnx_read_param(path.encode(), kind, name.encode(), None, 0)
    if n < 0:
        _check(int(n))
    out = np.empty(int(n), np.float32)
    L.wdr_onnx_read_param(path.encode(), kind, name.encode(), _p(out, f32p), int(n))
    return out


def silero_probe(path):
    hp = np.zeros(20, np.int32)
    nt = C.c_int32(0)
    _check(load().wdr_silero_probe(path.encode(), _p(hp, i32p), C.byref(nt)))
    return dict(version=tuple(int(v) for v in hp[:3]), n_encoder_layers=int(hp[3]), encoder=[tuple(int(v) for v in hp[4 + 3 * i: 7 + 3 * i]) for i in range(4)],
                lstm_input=int(hp[16]), lstm_hidden=int(hp[17]), final_in=int(hp[18]), final_out=int(hp[19]), n_tensors=int(nt.value))


def lang_str(i):
    r = load().wdr_lang_str(int(i))
    return r.decode() if r else None


def lang_id(s):
    return load().wdr_lang_id(s.encode())


SIZE_MAX = (1 << 64) - 1


class EmbeddingManager:
    """wdr_spk: pyannote_rs::EmbeddingManager (reference src/transcribe.rs:342, 480-492).  Pure host logic."""

    def __init__(self, max_speakers=SIZE_MAX):
        self.max_speakers = max_speakers
        self._h = load().wdr_spk_init(max_speakers)

    def close(self):
        if self._h:
            load().wdr_spk_free(self._h)
            self._h = None

    __del__ = close

    def count(self):
        return load().wdr_spk_count(self._h)

    def search_speaker(self, emb, threshold):
        e = _np(emb, np.float32)
        r = _check(load().wdr_spk_search(self._h, _p(e, f32p), len(e), float(threshold)))
        return r if r > 0 else None

    def get_best_speaker_match(self, emb):
        e = _np(emb, np.float32)
        return _check(load().wdr_spk_best_match(self._h, _p(e, f32p), len(e)))

    def assign_batch(self, emb, threshold):
        """The crate's policy over n embeddings in order (wdr_spk_assign_batch): labels [n] (0 = "?")."""
        e = _np(emb, np.float32)
        lab = np.zeros(e.shape[0], np.int32)
        if e.shape[0]:
            _check(load().wdr_spk_assign_batch(self._h, _p(e, f32p), e.shape[0], e.shape[1], threshold, _p(lab, i32p)))
        return lab

    def assign(self, emb, threshold):
        """The crate's policy (src/transcribe.rs:482-492): id, or None for "?"."""
        if self.count() == self.max_speakers:
            return self.get_best_speaker_match(emb)
        return self.search_speaker(emb, threshold)


def cosine_matrix(emb):
    e = _np(emb, np.float32)
    S = np.empty((e.shape[0], e.shape[0]), np.float32)
    _check(load().wdr_cosine_matrix(_p(e, f32p), e.shape[0], e.shape[1], _p(S, f32p)))
    return S


def cluster_leader(S, threshold, max_speakers=SIZE_MAX):
    s = _np(S, np.float32)
    labels = np.empty(s.shape[0], np.int32)
    _check(load().wdr_cluster_leader(_p(s, f32p), s.shape[0], float(threshold), max_speakers, _p(labels, i32p)))
    return labels


def cluster_agglomerative(S, threshold):
    s = _np(S, np.float32)
    labels = np.empty(s.shape[0], np.int32)
    _check(load().wdr_cluster_agglomerative(_p(s, f32p), s.shape[0], float(threshold), _p(labels, i32p)))
    return labels


def vad_default_params(**kw):
    p = load().wdr_vad_default_params()
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def _collect_segments(h):
    L = load()
    if not h:
        raise WdrError(-3, L.wdr_last_error().decode())
    try:
        return [(L.wdr_vad_segments_get_segment_t0(h, i), L.wdr_vad_segments_get_segment_t1(h, i)) for i in range(L.wdr_vad_segments_n(h))]
    finally:
        L.wdr_vad_free_segments(h)


def vad_segments_from_probs(probs, params=None):
    """whisper_vad_segments_from_probs on a caller-supplied probability array (host logic; no GPU needed) -> [(t0_cs, t1_cs)]."""
    pr = _np(probs, np.float32)
    return _collect_segments(load().wdr_vad_segments_from_probs_array(_p(pr, f32p), len(pr), params if params is not None else vad_default_params()))


class VadContext:
    """wdr_vad: WhisperVadContext (reference src/vad.rs:15-18)."""

    def __init__(self, seed=1234, gpu_device=0, path=None):
        L = load()
        p = L.wdr_vad_default_context_params()
        p.seed = seed
        p.gpu_device = gpu_device
        self._h = L.wdr_vad_init_from_file_with_params(path.encode() if path else None, p)  # path: ggml-silero-v5.1.2.bin
        if not self._h:
            raise WdrError(WDR_ERR_NO_DEVICE if device_count() == 0 else -3, L.wdr_last_error().decode())

    def close(self):
        if self._h:
            load().wdr_vad_free(self._h)
            self._h = None

    __del__ = close

    def detect_speech(self, pcm_f32):
        x = _np(pcm_f32, np.float32)
        _check(load().wdr_vad_detect_speech(self._h, _p(x, f32p), len(x)))
        n = load().wdr_vad_n_probs(self._h)
        return np.ctypeslib.as_array(load().wdr_vad_probs(self._h), shape=(n,)).copy() if n else np.zeros(0, np.float32)

    def detect_speech_batch(self, pcm_i16, offsets, n_samples):
        x = _np(pcm_i16, np.int16).reshape(-1)
        off, n = _np(offsets, np.int64), _np(n_samples, np.int32)
        nf = int(((n.astype(np.int64) + 511) // 512).sum())
        probs = np.empty(max(nf, 1), np.float32)
        foff = np.empty(len(n) + 1, np.int64)
        _check(load().wdr_vad_detect_speech_batch_i16(self._h, _p(x, i16p), _p(off, i64p), _p(n, i32p), len(n), _p(probs, f32p), _p(foff, i64p)))
        return [probs[foff[i]:foff[i + 1]].copy() for i in range(len(n))]

    def segments_from_samples(self, pcm_f32, params=None):
        """vad.segments_from_samples(params, &samples) -> [(start_cs, end_cs)] (reference src/vad.rs:31)."""
        x = _np(pcm_f32, np.float32)
        return _collect_segments(load().wdr_vad_segments_from_samples(self._h, params if params is not None else vad_default_params(), _p(x, f32p), len(x)))


def _collect_seg_result(h, with_samples):
    L = load()
    if not h:
        raise WdrError(-3, L.wdr_last_error().decode())
    try:
        out = []
        for i in range(L.wdr_seg_result_n(h)):
            i1 = C.c_int64(0)
            i0 = L.wdr_seg_result_sample_range(h, i, C.byref(i1))
            d = dict(start=L.wdr_seg_result_start(h, i), end=L.wdr_seg_result_end(h, i), i0=int(i0), i1=int(i1.value))
            if with_samples:
                cnt = C.c_int64(0)
                p = L.wdr_seg_result_samples(h, i, C.byref(cnt))
                d["samples"] = np.ctypeslib.as_array(p, shape=(cnt.value,)).copy() if cnt.value > 0 else np.zeros(0, np.int16)
            out.append(d)
        return out
    finally:
        L.wdr_seg_result_free(h)


def seg_segments_from_scores(scores, n_samples_padded):
    """pyannote-rs' speech state machine on scores[n_windows, 589, 7] (host logic; no GPU needed)."""
    sc = _np(scores, np.float32)
    return _collect_seg_result(load().wdr_seg_segments_from_scores(_p(sc, f32p), sc.shape[0], int(n_samples_padded)), False)


class Segmenter:
    """wdr_seg: the segmentation-3.0 session pyannote_rs::get_segments opens (reference src/engine.rs:117)."""

    def __init__(self, seed=1234, device=0, path=None):
        self._h = load().wdr_seg_init(path.encode() if path else None, seed, device)  # path: segmentation-3.0.onnx
        if not self._h:
            raise WdrError(WDR_ERR_NO_DEVICE if device_count() == 0 else -3, load().wdr_last_error().decode())

    def close(self):
        if self._h:
            load().wdr_seg_free(self._h)
            self._h = None

    __del__ = close

    def scores(self, pcm_i16):
        x = _np(pcm_i16, np.int16)
        W = load().wdr_seg_n_windows(len(x))
        out = np.empty((W, 589, 7), np.float32)
        if W:
            _check(load().wdr_seg_scores_i16(self._h, _p(x, i16p), len(x), _p(out, f32p)))
        return out

    def get_segments(self, pcm_i16):
        """pyannote_rs::get_segments(&samples, 16000, model) -> [dict(start, end, samples)]."""
        x = _np(pcm_i16, np.int16)
        return _collect_seg_result(load().wdr_seg_get_segments(self._h, _p(x, i16p), len(x)), True)


class EmbeddingExtractor:
    """wdr_emb: pyannote_rs::EmbeddingExtractor (reference src/transcribe.rs:343, 466-467) — WeSpeaker ResNet34, 256-d."""

    def __init__(self, seed=1234, device=0, path=None):
        self._h = load().wdr_emb_init(path.encode() if path else None, seed, device)  # path: a WeSpeaker ResNet34 .onnx export
        if not self._h:
            raise WdrError(WDR_ERR_NO_DEVICE if device_count() == 0 else -3, load().wdr_last_error().decode())
        self.dim = load().wdr_emb_dim(self._h)

    def close(self):
        if self._h:
            load().wdr_emb_free(self._h)
            self._h = None

    __del__ = close

    def compute(self, pcm_i16):
        """compute(&samples) -> embedding [256]; raises WdrError(WDR_ERR_TOO_SHORT) when fbank yields no frame."""
        x = _np(pcm_i16, np.int16)
        out = np.empty(self.dim, np.float32)
        _check(load().wdr_emb_compute_i16(self._h, _p(x, i16p), len(x), _p(out, f32p)))
        return out

    def compute_batch(self, pcm_i16, seg_offset):
        """Segments pcm[seg_offset[s]:seg_offset[s+1]] -> (emb [n, 256], status [n] (0 or WDR_ERR_TOO_SHORT))."""
        x = _np(pcm_i16, np.int16)
        so = _np(seg_offset, np.int64)
        n = len(so) - 1
        out = np.zeros((n, self.dim), np.float32)
        status = np.zeros(n, np.int32)
        _check(load().wdr_emb_compute_batch_i16(self._h, _p(x, i16p), _p(so, i64p), n, _p(out, f32p), _p(status, i32p)))
        return out, status

    def compute_batch_dev(self, pcm_dev_ptr, seg_offset, out_dev_ptr, stream=None):
        so = _np(seg_offset, np.int64)
        n = len(so) - 1
        status = np.zeros(n, np.int32)
        _check(load().wdr_emb_compute_batch_i16_dev(self._h, pcm_dev_ptr, _p(so, i64p), n, out_dev_ptr, _p(status, i32p), stream))
        return status

    def last_flops(self):
        return float(load().wdr_emb_last_flops(self._h))

    def profile(self, on=True):
        _check(load().wdr_emb_profile(self._h, int(on)))

    def last_kernel_ms(self):
        """(tcgen05 GEMM ms, im2col gather ms) of the last compute call (profiling on)."""
        a, b = C.c_double(0), C.c_double(0)
        _check(load().wdr_emb_last_kernel_ms(self._h, C.byref(a), C.byref(b)))
        return float(a.value), float(b.value)


DIST_ID_BYTES = 128


def dist_available():
    return bool(load().wdr_dist_available())


def dist_unique_id():
    """ncclGetUniqueId through the C ABI (rank 0): 128 bytes to hand to every rank over the host's own channel."""
    buf = (C.c_uint8 * DIST_ID_BYTES)()
    _check(load().wdr_dist_get_unique_id(buf))
    return bytes(buf)


class Dist:
    """wdr_dist: this rank's handle on the path's one exchange (the all-gather of speaker embeddings), NCCL inside the library."""

    def __init__(self, unique_id, n_ranks, rank, device):
        buf = (C.c_uint8 * DIST_ID_BYTES).from_buffer_copy(unique_id) if unique_id is not None else None
        self._h = load().wdr_dist_init(buf, n_ranks, rank, device)
        if not self._h:
            raise WdrError(-3, load().wdr_last_error().decode())
        self.n_ranks, self.rank = n_ranks, rank

    def close(self):
        if self._h:
            load().wdr_dist_free(self._h)
            self._h = None

    __del__ = close

    def allgather_embeddings(self, emb_local, max_rows, normalize=False, dim=None):
        """emb_local [n, D] host -> ([sum n_r, D] in rank order, counts [n_ranks]).  A rank without rows passes dim = D."""
        e = _np(emb_local, np.float32)
        D = int(dim) if dim is not None else int(e.shape[1])
        e = e.reshape(-1, D)
        out = np.empty((max_rows, D), np.float32)
        counts = np.zeros(self.n_ranks, np.int32)
        n = _check(load().wdr_allgather_embeddings(self._h, _p(e, f32p), e.shape[0], D, int(normalize), _p(out, f32p), max_rows, _p(counts, i32p)))
        return out[:n], counts

    def allgather_embeddings_dev(self, emb_dev_ptr, n_local, D, out_dev_ptr, out_cap_rows, normalize=False, stream=None):
        counts = np.zeros(self.n_ranks, np.int32)
        n = _check(load().wdr_allgather_embeddings_dev(self._h, emb_dev_ptr, n_local, D, int(normalize), out_dev_ptr, out_cap_rows, _p(counts, i32p), stream))
        return n, counts
