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


def diarize(segmenter, extractor, int_samples, threshold=0.5, max_speakers=capi.SIZE_MAX, mode="leader"):
    """The crate's diarize flow around the boundary, batched: pyannote_rs::get_segments (reference src/engine.rs:117-122) ->
    EmbeddingExtractor::compute per segment (src/transcribe.rs:466-467) -> speaker id per segment.

    mode "leader" = the reference's policy (src/transcribe.rs:480-496): EmbeddingManager in segment order — strict `> threshold`
    joins the best stored speaker, else a new speaker while fewer than max_speakers exist, else (cap reached) the best match;
    a segment too short for one fbank frame gets "?".  Computed here as the ordered scan of the pairwise cosine matrix
    (wdr_cosine_matrix + wdr_cluster_leader), which is the same function of the similarities (tests/test_oracle_cluster.py).
    mode "agglomerative" = the north-star's average-linkage clustering of the same matrix.
    Returns [dict(start, end, speaker)] with speaker a str as the crate renders it ("1", "2", ... or "?")."""
    int_samples = np.asarray(int_samples, np.int16)
    segs = segmenter.get_segments(int_samples)
    if not segs:
        return []
    off = np.zeros(len(segs) + 1, np.int64)
    for i, s in enumerate(segs):
        off[i + 1] = off[i] + len(s["samples"])
    pcm = np.concatenate([s["samples"] for s in segs]).astype(np.int16) if off[-1] else np.zeros(0, np.int16)
    emb, status = extractor.compute_batch(pcm, off)
    ok = np.flatnonzero(status == 0)
    speakers = ["?"] * len(segs)
    if len(ok):
        S = capi.cosine_matrix(emb[ok])
        labels = capi.cluster_leader(S, threshold, max_speakers) if mode == "leader" else capi.cluster_agglomerative(S, threshold)
        for i, l in zip(ok, labels):
            speakers[i] = str(int(l)) if l > 0 else "?"
    return [dict(start=s["start"], end=s["end"], speaker=sp) for s, sp in zip(segs, speakers)]
