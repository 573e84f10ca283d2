"""GPU suite: the crate-shaped flow `Engine::transcribe_audio` -> `run_transcription_pipeline` (reference src/engine.rs:65-200,
src/transcribe.rs:323-535) driven through the C ABI by whisper-diarize-rs_b200/host.py, in its three modes: whole file (one
SpeechSegment, whisper_full's sequential loop), Silero VAD segments, pyannote segmentation + speaker ids."""
import numpy as np
import pytest

from conftest import synth_audio

pytestmark = pytest.mark.gpu


def _check_segments(segs, t_max):
    last_end = 0.0
    for s in segs:
        assert 0.0 <= s["start"] <= s["end"] <= t_max + 1e-6
        assert s["start"] >= last_end - 1e-9, "segments must not overlap after clipping"
        last_end = s["end"]
        if s["words"]:
            assert s["start"] == s["words"][0]["start"]
            for w in s["words"]:  # (a clipped segment's last word may end before it starts: the crate clips only the end, :447-455)
                assert not w["text"].startswith("[_")


def test_transcribe_audio_three_modes(wdr):
    from hostmirror import host as H
    pcm = synth_audio(71, 42.0, n_speakers=2)
    ctx = wdr.Context("tiny.en", seed=1234, enable_dtw=True)
    st = ctx.create_state()
    # 1. whole file: one SpeechSegment of 42 s -> sequential mode inside state.full
    segs, lang, mask = H.transcribe_audio(st, pcm)
    assert lang == "en" and mask is None and len(segs) >= 2
    _check_segments(segs, 42.0 + 30.0)
    direct = st.full(pcm)
    assert [s["text"] for s in segs] == [d["text"].lstrip() for d in direct]
    # the consumer (reference src/formatting.rs via formatting.py): cues from the library's token spans
    cues = H.format_cues(segs, lang, None, dict(max_lines=2))
    assert cues and all(c["text"].count("\n") <= 1 and "[_" not in c["text"] and c["words"] for c in cues)
    assert [c["start"] for c in cues] == sorted(c["start"] for c in cues)
    n_chars = sum(len(w["text"]) for c in cues for w in c["words"])
    assert n_chars == sum(len(w["text"].strip()) for s in segs for w in (s["words"] or []))  # nothing printable is lost or invented
    # 2. VAD segments
    vad = wdr.VadContext(seed=1234)
    segs_v, _, mask_v = H.transcribe_audio(st, pcm, enable_vad=True, vad=vad)
    assert mask_v is not None
    assert H.format_cues(segs_v, "en", mask_v) is not None
    _, speech = H.vad_get_segments(vad, pcm)
    assert len(segs_v) <= sum(1 for _ in speech) * 2 + 2
    for s in segs_v:
        assert any(sp["start"] - 1e-6 <= s["start"] for sp in speech)
    _check_segments(segs_v, 42.0 + 30.0)
    # 3. diarize: pyannote segments + speaker ids ("1", "2", ... or "?")
    seg_m = wdr.Segmenter(seed=1234)
    emb = wdr.EmbeddingExtractor(seed=1234)
    segs_d, _, _ = H.transcribe_audio(st, pcm, enable_diarize=True, segmenter=seg_m, extractor=emb, max_speakers=2, carry_prompt=False)
    assert segs_d and all(s["speaker_id"] in ("1", "2", "?") for s in segs_d)
    # the speaker of a speech segment equals the manual EmbeddingManager scan over the same segments
    speech_d = seg_m.get_segments(pcm)
    mgr = wdr.EmbeddingManager(2)
    want = {}
    for sp in speech_d:
        got = st.full(sp["samples"])
        if not got:
            continue
        try:
            sid = mgr.assign(emb.compute(sp["samples"]), 0.5)
            want[round(sp["start"], 6)] = str(sid) if sid else "?"
        except wdr.WdrError:
            want[round(sp["start"], 6)] = "?"
    mgr.close()
    assert sorted(set(s["speaker_id"] for s in segs_d)) == sorted(set(want.values()))
    for m in (vad, seg_m, emb):
        m.close()
    st.close()
    ctx.close()


def test_sharded_pipeline_equals_sequential(wdr):
    """run_transcription_pipeline_sharded (all speech segments through ONE wdr_full_batch_i16 / ONE wdr_emb_compute_batch_i16 call)
    returns exactly what the crate-shaped loop returns without prompt carry (one state.full + one EmbeddingExtractor::compute per
    segment): same segments, texts, word spans to the last bit, same speaker ids — for pyannote segments (diarize) and for Silero VAD
    segments, including a speech segment longer than 30 s (which keeps whisper_full's sequential seek loop)."""
    from hostmirror import host as H
    pcm = synth_audio(73, 48.0, n_speakers=3)
    ctx = wdr.Context("tiny.en", seed=1234, enable_dtw=True)
    st = ctx.create_state()
    seg_m = wdr.Segmenter(seed=1234)
    emb = wdr.EmbeddingExtractor(seed=1234)
    speech = [dict(start=s["start"], end=s["end"], samples=s["samples"]) for s in seg_m.get_segments(pcm)]
    assert len(speech) >= 3
    speech.append(dict(start=50.0, end=50.0 + 36.0, samples=synth_audio(74, 36.0)))  # > 30 s: sequential mode inside the sharded call
    a, lang_a = H.run_transcription_pipeline(st, speech, None, emb, 0.5, 2, carry_prompt=False)
    b, lang_b = H.run_transcription_pipeline_sharded(st, speech, None, emb, 0.5, 2, batch=5)  # several batches
    assert lang_a == lang_b == "en" and len(a) == len(b) and len(a) >= 3
    for x, y in zip(a, b):
        assert (x["start"], x["end"], x["text"], x["speaker_id"]) == (y["start"], y["end"], y["text"], y["speaker_id"])
        assert x["words"] == y["words"]
    vad = wdr.VadContext(seed=1234)
    _, speech_v = H.vad_get_segments(vad, pcm)
    a, _ = H.run_transcription_pipeline(st, speech_v, carry_prompt=False)
    b, _ = H.run_transcription_pipeline_sharded(st, speech_v)
    assert len(a) == len(b)
    for x, y in zip(a, b):
        assert (x["start"], x["end"], x["text"]) == (y["start"], y["end"], y["text"]) and x["words"] == y["words"]
    for m in (vad, seg_m, emb):
        m.close()
    st.close()
    ctx.close()
