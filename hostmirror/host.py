"""Host-side mirror of the reference's own Rust logic that sits directly around the C-ABI boundary (same names, argument
meaning and error behaviour), so that the parity tests read like the crate's call sites.  No Rust toolchain exists in the
build image; in production this logic stays in the crate (reference src/vad.rs, src/transcribe.rs) and only the FFI target
changes."""
import numpy as np

import wdr_b200 as capi  # the ctypes binding over include/wdr.h (the product boundary)

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
    def idx(t):  # ((t as f32 * SR).round()).clamp(0.0, n_f32) as usize — f32::round is half away from zero; done in f64, where
        x = float(F(F(t) * SR))  # x + 0.5 is exact for every f32 x (floor(x + 0.5) in f32 is off by one for odd x above 2^23)
        r = np.floor(x + 0.5) if x >= 0.0 else -np.floor(-x + 0.5)
        return int(min(max(r, 0.0), float(n_f32)))

    for s, e in merged:                                                              # :66-81
        si, ei = idx(s), idx(e)
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


# ---------------------------------------------------------------------------------------------------------------------
# reference src/utils.rs and src/transcribe.rs logic that consumes the boundary's results (pure integer / f64 arithmetic)
# ---------------------------------------------------------------------------------------------------------------------
def calculate_dtw_mem_size(num_samples):
    """reference src/utils.rs:3-49: DTW working-set estimate handed to DtwParameters (accepted by wdr_context_params)."""
    num_frames = (num_samples + 159) // 160
    band = 96 if num_frames <= 15000 else 128 if num_frames <= 45000 else 160
    total = 24 * 1024 * 1024 + num_frames * band * 4 * 4 + num_frames * 4
    clamped = min(max(total, 24 * 1024 * 1024), 768 * 1024 * 1024)
    align = 8 * 1024 * 1024
    return (clamped + align - 1) & ~(align - 1)


def cs_to_s(cs):
    """reference src/utils.rs:57-59."""
    return float(cs) * 0.01


def is_whole_control_token(s):
    """reference src/transcribe.rs:206-212: "[_BEG_]", "[_TT_320]", ... (how whisper.cpp / libwdr_b200 print special tokens)."""
    t = s.strip("\0").strip()
    if not (t.startswith("[_") and t.endswith("]")):
        return False
    inner = t[2:-1]
    return bool(inner) and all(("A" <= c <= "Z") or ("0" <= c <= "9") or c == "_" for c in inner)


def strip_embedded_control_markers(s):
    """reference src/transcribe.rs:215-240."""
    out, i = [], 0
    while i < len(s):
        if i + 1 < len(s) and s[i] == "[" and s[i + 1] == "_":
            j = i + 2
            while j < len(s) and s[j] != "]":
                j += 1
            if j < len(s) and is_whole_control_token(s[i:j + 1]):
                i = j + 1
                continue
        out.append(s[i])
        i += 1
    return "".join(out)


def get_token_timestamps(token_texts, token_data):
    """reference src/transcribe.rs:242-320 on one segment: token_texts[i] = to_str_lossy(), token_data[i] has .p .t0 .t1 .t_dtw
    (centiseconds; t_dtw < 0 = no DTW anchor).  Returns [dict(text, start, end, probability)] in seconds relative to the buffer."""
    toks = []
    for raw, td in zip(token_texts, token_data):
        if is_whole_control_token(raw):
            continue
        clean = strip_embedded_control_markers(raw)
        if not clean.strip("\0").strip():
            continue
        toks.append(dict(text=clean, p=float(td.p), t0=cs_to_s(td.t0), t1=cs_to_s(td.t1), anchor=cs_to_s(td.t_dtw) if td.t_dtw >= 0 else None))
    spans = []
    for i, t in enumerate(toks):
        a_prev = toks[i - 1]["anchor"] if i > 0 else None
        a_here = t["anchor"]
        a_next = toks[i + 1]["anchor"] if i + 1 < len(toks) else None
        start = 0.5 * (a_prev + a_here) if (a_prev is not None and a_here is not None) else t["t0"]
        end = 0.5 * (a_here + a_next) if (a_here is not None and a_next is not None) else t["t1"]
        spans.append(dict(text=t["text"], start=start, end=end, probability=t["p"]))
    return spans


def interpolate_word_timestamps(line, start, end):
    """reference src/transcribe.rs:171-203 (translate task: token timings no longer align with the words)."""
    dur = max(end - start, 0.0)
    if dur <= 0.0:
        return []
    tokens = [t for t in line.split() if t.strip("\0").strip()]
    if not tokens:
        return []
    weights = [max(sum(1 for c in t if c.isalnum()), 1) for t in tokens]
    total = sum(weights)
    out, acc = [], 0
    for i, tok in enumerate(tokens):
        t0 = start + (acc / total) * dur
        t1 = end if i + 1 == len(tokens) else start + ((acc + weights[i]) / total) * dur
        acc += weights[i]
        out.append(dict(text=tok, start=t0, end=t1, probability=None))
    return out


def assemble_segments(state_segments, base_offset, segments_out, translated=False):
    """reference src/transcribe.rs:397-459 for the segments of one state.full call (State.segments() dicts): trimmed text, absolute
    word timestamps (base_offset = SpeechSegment.start + user offset), bounds from the first / last word, and the clipping of the
    previous segment's end (and its last word) to the new segment's start.  Appends to segments_out and returns it."""
    for seg in state_segments:
        text = seg["text"].lstrip()
        approx_start = base_offset + cs_to_s(seg["t0"])
        approx_end = base_offset + cs_to_s(seg["t1"])
        if translated:
            words = interpolate_word_timestamps(text, approx_start, approx_end)
        else:
            words = get_token_timestamps(seg["token_text"], seg["tokens"])
            for w in words:
                w["start"] += base_offset
                w["end"] += base_offset
        seg_start = words[0]["start"] if words else approx_start
        seg_end = words[-1]["end"] if words else approx_end
        if segments_out:
            last = segments_out[-1]
            if last["end"] > seg_start:
                last["end"] = seg_start
            if last["words"]:
                if last["words"][-1]["end"] > last["end"]:
                    last["words"][-1]["end"] = last["end"]
        segments_out.append(dict(start=seg_start, end=seg_end, text=text, words=words or None, speaker_id=None))
    return segments_out


# ---------------------------------------------------------------------------------------------------------------------
# Engine::transcribe_audio / run_transcription_pipeline (reference src/engine.rs:65-200, src/transcribe.rs:323-535)
# ---------------------------------------------------------------------------------------------------------------------
def run_transcription_pipeline(state, speech_segments, params=None, extractor=None, threshold=0.5, max_speakers=capi.SIZE_MAX,
                               user_offset=0.0, carry_prompt=True, translated=False):
    """reference src/transcribe.rs:323-535 over the C ABI: for every SpeechSegment dict(start, end, samples int16) in order —
    `state.full` on its samples (one buffer; > 30 s runs whisper_full's own seek loop), segments -> words (get_token_timestamps),
    absolute times (base_offset = segment.start + user offset), overlap clipping against the previous segment, and — with an
    EmbeddingExtractor — the speaker of the SPEECH segment's samples through an EmbeddingManager (:461-497; "?" when the embedding
    fails).  The crate feeds the previous segment's (left-trimmed) text back as `initial_prompt` (:383-386, :502): the library
    tokenises it with the context's vocabulary as whisper_full does (wdr_tokenize).
    Returns (segments [dict(start, end, text, words, speaker_id)], detected_lang)."""
    mgr = capi.EmbeddingManager(max_speakers) if extractor is not None else None
    out, previous_text, detected = [], None, None
    # `params.set_initial_prompt` mutates the params the loop clones: once set, a prompt stays until replaced — and a prompt the
    # caller set (advanced.init_prompt, :74-76) stays until the first non-empty segment text replaces it
    sticky = params.initial_prompt if params is not None and params.initial_prompt else None
    for sp in speech_segments:
        p = params if params is not None else state.full_params()
        p.prompt_tokens = None
        p.prompt_n_tokens = 0
        if carry_prompt and previous_text is not None:                                   # :383-386
            sticky = previous_text.encode()
        keep = sticky                                                                    # kept alive for the call (borrowed string)
        p.initial_prompt = keep
        segs = state.full(np.asarray(sp["samples"], np.int16), p)                      # :389
        del keep
        if detected is None:
            detected = capi.lang_str(state.lang_id())                                 # :391-395
        n_before = len(out)
        assemble_segments(segs, sp["start"] + user_offset, out, translated)          # :397-459
        speaker = None
        if extractor is not None and segs:                                            # :461-497
            try:
                emb = extractor.compute(sp["samples"])
                sid = mgr.assign(emb, threshold)
                speaker = str(sid) if sid else "?"
            except capi.WdrError:
                speaker = "?"
        for s in out[n_before:]:
            s["speaker_id"] = speaker
            previous_text = s["text"] if s["text"].strip() else None                    # :502 (updated per whisper segment)
    if mgr is not None:
        mgr.close()
    return out, detected


def run_transcription_pipeline_sharded(state, speech_segments, params=None, extractor=None, threshold=0.5, max_speakers=capi.SIZE_MAX,
                                       user_offset=0.0, translated=False, batch=128):
    """The same pipeline at throughput (SURVEY §0.4's independent-chunk mode, §8e): every SpeechSegment is its own `state.full`
    buffer, so — once the previous segment's text is no longer fed back as a prompt — the segments are independent and go through
    the library `batch` at a time: ONE wdr_full_batch_i16 call per `batch` segments of <= 30 s (mel, encoder, decode and DTW of all
    of them as one batch), ONE wdr_emb_compute_batch_i16 call for all speaker embeddings, then the crate's host logic in segment
    order (assemble_segments, overlap clipping, EmbeddingManager).  Segments longer than 30 s keep whisper_full's sequential seek loop
    (one `state.full` each).  Output == run_transcription_pipeline(..., carry_prompt=False) on the same input."""
    speech_segments = list(speech_segments)
    per_seg = [None] * len(speech_segments)
    short = [i for i, sp in enumerate(speech_segments) if len(sp["samples"]) <= 480000]
    p = params if params is not None else state.full_params()
    p.prompt_tokens = None
    p.prompt_n_tokens = 0
    detected = None
    for o in range(0, len(short), batch):
        ids = short[o: o + batch]
        pcm = np.zeros((len(ids), 480000), np.int16)
        nv = np.zeros(len(ids), np.int32)
        for r, i in enumerate(ids):
            x = np.asarray(speech_segments[i]["samples"], np.int16)
            pcm[r, : len(x)] = x
            nv[r] = len(x)
        segs = state.full_batch(pcm, nv, p)
        if detected is None and ids and ids[0] == 0:
            detected = capi.lang_str(state.chunk_lang_id(0))
        for r, i in enumerate(ids):
            per_seg[i] = [sg for sg in segs if sg["chunk"] == r]
    for i, sp in enumerate(speech_segments):
        if per_seg[i] is None:
            per_seg[i] = state.full(np.asarray(sp["samples"], np.int16), p)
            if detected is None and i == 0:
                detected = capi.lang_str(state.lang_id())
    emb = status = None
    if extractor is not None and speech_segments:
        off = np.zeros(len(speech_segments) + 1, np.int64)
        for i, sp in enumerate(speech_segments):
            off[i + 1] = off[i] + len(sp["samples"])
        cat = np.concatenate([np.asarray(sp["samples"], np.int16) for sp in speech_segments]) if off[-1] else np.zeros(0, np.int16)
        emb, status = extractor.compute_batch(cat, off)
    mgr = capi.EmbeddingManager(max_speakers) if extractor is not None else None
    out = []
    for i, sp in enumerate(speech_segments):
        segs = per_seg[i]
        n_before = len(out)
        assemble_segments(segs, sp["start"] + user_offset, out, translated)
        speaker = None
        if extractor is not None and segs:
            if status[i] == 0:
                sid = mgr.assign(emb[i], threshold)
                speaker = str(sid) if sid else "?"
            else:
                speaker = "?"
        for sgm in out[n_before:]:
            sgm["speaker_id"] = speaker
    if mgr is not None:
        mgr.close()
    if detected is None:
        detected = capi.lang_str(state.lang_id())
    return out, detected


def format_cues(segments, lang, vad_mask=None, formatting_overrides=None):
    """The tail of Engine::transcribe_audio (reference src/engine.rs:189-198): PostProcessConfig::for_language(effective_lang) +
    overrides, process_segments with the VAD mask as the silence oracle (formatting.py restates src/formatting.rs)."""
    from . import formatting as F
    cfg = F.config_for_language(lang or "en", formatting_overrides)
    return F.process_segments(segments, cfg, F.VadMaskOracle(vad_mask) if vad_mask is not None else None)


def transcribe_audio(state, int_samples, enable_vad=False, enable_diarize=False, vad=None, segmenter=None, extractor=None,
                     threshold=0.5, max_speakers=None, params=None, carry_prompt=True):
    """reference src/engine.rs:65-200 after `read_wav`: choose the speech segments (pyannote segmentation when diarizing, Silero VAD
    segments when enabled, else ONE segment holding the whole file), run the pipeline, return (segments, detected_lang, vad_mask);
    `format_cues` is the formatting tail (translation over HTTP stays in the crate)."""
    int_samples = np.asarray(int_samples, np.int16)
    mask = None
    if enable_diarize:                                                                # :88-122
        cap = capi.SIZE_MAX if not max_speakers else max_speakers                     # Some(0) | None => usize::MAX
        speech = [dict(start=s["start"], end=s["end"], samples=s["samples"]) for s in segmenter.get_segments(int_samples)]
        segs, lang = run_transcription_pipeline(state, speech, params, extractor, threshold, cap, carry_prompt=carry_prompt)
    elif enable_vad:                                                                  # :123-140
        mask, speech = vad_get_segments(vad, int_samples)
        segs, lang = run_transcription_pipeline(state, speech, params, carry_prompt=carry_prompt)
    else:                                                                             # :141-147
        speech = [dict(start=0.0, end=len(int_samples) / 16000.0, samples=int_samples)]
        segs, lang = run_transcription_pipeline(state, speech, params, carry_prompt=carry_prompt)
    return segs, lang, mask
