"""Host-side mirror of the reference's own Rust logic that sits directly around the C-ABI boundary (same names, argument
meaning and error behaviour), so that the parity tests read like the crate's call sites.  No Rust toolchain exists in the
build image; in production this logic stays in the crate (reference src/vad.rs, src/transcribe.rs) and only the FFI target
changes."""
import numpy as np

from . import capi

F = np.float32


def convert_integer_to_float_audio(int_samples):
    """whisper_rs::convert_integer_to_float_audio (reference src/vad.rs:11-12): x / 32768."""
    return np.asarray(int_samples, np.int16).astype(F) / F(32768.0)


def vad_get_segments(vad, int_samples):
    """reference src/vad.rs:6-85 `get_segments(vad_model, int_samples)` with an already created VadContext.
    Returns (mask, merged_segments): mask = [(start_s, end_s)] raw speech ranges; merged_segments = [dict(start, end, samples)]."""
    int_samples = np.asarray(int_samples, np.int16)
    samples = convert_integer_to_float_audio(int_samples)
    vadp = capi.vad_default_params(min_silence_duration_ms=100)                      # src/vad.rs:21-22
    segs = vad.segments_from_samples(samples, vadp)                                  # src/vad.rs:31
    return vad_mask_and_merge(segs, int_samples)


def vad_mask_and_merge(segs_cs, int_samples):
    """reference src/vad.rs:33-82 on the VAD's [(start_cs, end_cs)] list."""
    n = len(int_samples)
    SR = F(16000.0)
    n_f32 = F(n)
    mask = [(float(F(s)) / 100.0, float(F(e)) / 100.0) for s, e in segs_cs]          # :40-43
    mask = [(s, e) for s, e in mask if e > s]
    mask.sort(key=lambda t: t[0])                                                    # :46
    merged = []
    for s, e in mask:                                                                # :49-63, MERGE_GAP_S = 0.200
        if merged and s - merged[-1][1] < 0.200:
            merged[-1][1] = max(e, merged[-1][1])
        else:
            merged.append([s, e])
    out = []
    for s, e in merged:                                                              # :66-81 (f32 arithmetic, round half away from zero)
        si = int(np.clip(np.floor(F(F(s) * SR) + F(0.5)), F(0.0), n_f32))
        ei = int(np.clip(np.floor(F(F(e) * SR) + F(0.5)), F(0.0), n_f32))
        seg = int_samples[si:ei] if ei > si else int_samples[:0]
        if e > s and len(seg):
            out.append(dict(start=s, end=e, samples=seg))
    return mask, out
