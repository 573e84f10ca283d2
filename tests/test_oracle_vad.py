"""CPU suite for row a8: known-answer tests for whisper_vad_segments_from_probs (oracle) and bit-exact agreement of the
library's host-side segmenter (wdr_vad_segments_from_probs_array runs without a GPU) on hand-built and random probability
tracks; the crate's own mask/merge/slice logic (reference src/vad.rs:33-82)."""
import numpy as np


def test_segments_from_probs_kat():
    from oracle import vad as V
    p = np.zeros(200, np.float32)
    p[10:40] = 0.9     # 30 frames = 0.96 s of speech starting at sample 5120
    p[60:63] = 0.9     # 3 frames = 96 ms < min_speech 250 ms: dropped
    p[100:150] = 0.6
    p[120:122] = 0.4   # dip between neg_threshold (0.35) and threshold: speech continues
    segs = V.segments_from_probs(p, dict(min_silence_duration_ms=100))
    # segment 1: start 10*512 - pad(480) = 4640 -> 29 cs; end: silence first seen at frame 40 (20480) + pad 480 = 20960 -> 131 cs
    assert segs[0] == (29.0, 131.0)
    assert len(segs) == 2
    s1 = V.samples_to_cs(100 * 512 - 480), V.samples_to_cs(150 * 512 + 480)
    assert segs[1] == (float(s1[0]), float(s1[1]))
    # all silence / all speech
    assert V.segments_from_probs(np.zeros(50, np.float32)) == []
    full = V.segments_from_probs(np.ones(50, np.float32))
    assert full == [(0.0, float(V.samples_to_cs(50 * 512)))]
    assert V.segments_from_probs(np.zeros(0, np.float32)) == []


def test_library_segmenter_is_bit_exact(wdr):
    from oracle import vad as V
    rng = np.random.default_rng(0)
    for trial in range(40):
        n = int(rng.integers(1, 400))
        # piecewise-constant tracks with noise: many threshold crossings, short gaps, short bursts
        p = np.repeat(rng.random(n // 7 + 1), 7)[:n].astype(np.float32)
        p = np.clip(p + 0.1 * rng.standard_normal(n).astype(np.float32), 0, 1).astype(np.float32)
        for kw in (dict(min_silence_duration_ms=100), dict(), dict(threshold=0.3, speech_pad_ms=100, min_speech_duration_ms=100),
                   dict(max_speech_duration_s=2.0, min_silence_duration_ms=100)):
            ref = V.segments_from_probs(p, kw)
            got = wdr.vad_segments_from_probs(p, wdr.vad_default_params(**kw))
            assert got == ref, (trial, kw)


def test_mask_merge_slice_matches_crate_logic(wdr):
    from oracle import vad as V
    pcm = (np.arange(160000) % 1000).astype(np.int16)
    segs = [(10.0, 50.0), (60.0, 90.0), (300.0, 310.0), (120.0, 119.0), (990.0, 1200.0)]  # cs; gap 0.1 s merges, inverted range drops
    m1, s1 = V.get_segments(segs, pcm)
    m2, s2 = __import__("hostmirror").host.vad_mask_and_merge(segs, pcm)
    assert m1 == m2 == [(0.1, 0.5), (0.6, 0.9), (3.0, 3.1), (9.9, 12.0)]
    assert [(a, b, len(c)) for a, b, c in s1] == [(d["start"], d["end"], len(d["samples"])) for d in s2]
    assert [(round(d["start"], 3), round(d["end"], 3), len(d["samples"])) for d in s2] == [(0.1, 0.9, 12800), (3.0, 3.1, 1600), (9.9, 12.0, 1600)]
    assert np.array_equal(s2[0]["samples"], pcm[1600:14400])
