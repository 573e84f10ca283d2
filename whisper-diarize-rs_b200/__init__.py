"""whisper-diarize-rs B200 hot path: Python host-side mirror over the C ABI (include/wdr.h).

The compute lives in csrc/libwdr_b200.so (hand-written CUDA for sm_100a).  This package is a thin
ctypes binding used by tests/ and bench.py; it never falls back to a CPU implementation: if the shared
library is missing `load()` raises, and on a box without a GPU every compute call raises WdrError with
WDR_ERR_NO_DEVICE.

The directory name contains a hyphen (it mirrors the reference crate's name), so import it with
    importlib.import_module("whisper-diarize-rs_b200")
or through the `wdr_b200` alias module at the repo root.
"""
from .capi import (  # noqa: F401
    WdrError, load, lib_path, build, version, device_count, launch_count,
    MelFrontend, log_mel, median_filter, dtw_cost, dtw, dtw_batch_dev, kaldi_fbank, fbank_frames,
    signal_energy, convert_integer_to_float_audio, resample_to_16k, sample_discrete, tokenize_with_vocab, mel_n_len, gemm_bf16_dev,
    Context, State, ContextParams, ModelDims, mel_filters, encoder_attention_dev, DTW_PRESETS,
    FullParams, TokenData, lang_str, lang_id, ggml_probe, onnx_probe, onnx_read_param, silero_probe, ONNX_PYANNET, ONNX_RESNET34,
    VadContext, VadParams, vad_default_params, vad_segments_from_probs,
    Segmenter, seg_segments_from_scores, EmbeddingExtractor,
    EmbeddingManager, Dist, dist_available, dist_unique_id, cosine_matrix, cluster_leader, cluster_agglomerative, SIZE_MAX,
)
from . import dist  # noqa: F401,E402
