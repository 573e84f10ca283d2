"""GPU parity for row a8: Silero-style frame scoring on the device against the numpy oracle (fp32 network, tolerance 2e-4 on
the probabilities), batched streams == single streams, and the full `get_segments` chain of reference src/vad.rs:6-85."""
import numpy as np
import pytest

from conftest import synth_audio

pytestmark = pytest.mark.gpu


def test_probs_match_oracle_and_batch_equals_single(wdr):
    from oracle import vad as V
    w = V.vad_weights(1234)
    vad = wdr.VadContext(seed=1234)
    pcms = [synth_audio(31, 6.0), synth_audio(32, 3.3)[:-77], np.zeros(2000, np.int16), synth_audio(33, 0.02)]
    singles = []
    for pcm in pcms:
        x = pcm.astype(np.float32) / np.float32(32768.0)
        got = vad.detect_speech(x)
        ref = V.silero_probs(x, w)
        assert got.shape == ref.shape == ((len(x) + 511) // 512,)
        assert np.abs(got - ref).max() < 2e-4, np.abs(got - ref).max()
        singles.append(got)
    cat = np.concatenate(pcms)
    n = np.array([len(p) for p in pcms], np.int32)
    off = np.concatenate([[0], np.cumsum(n)[:-1]]).astype(np.int64)
    batch = vad.detect_speech_batch(cat, off, n)
    for a, b in zip(batch, singles):
        assert np.array_equal(a, b)  # same kernels, same order of operations
    assert len(vad.detect_speech(np.zeros(0, np.float32))) == 0
    vad.close()


def test_get_segments_chain(wdr):
    from oracle import vad as V
    w = V.vad_weights(1234)
    vad = wdr.VadContext(seed=1234)
    pcm = synth_audio(41, 20.0, n_speakers=2)
    mask, segs = __import__("hostmirror").host.vad_get_segments(vad, pcm)
    x = pcm.astype(np.float32) / np.float32(32768.0)
    probs = vad.detect_speech(x)
    # the host segmenter applied to the device probabilities must equal the oracle segmenter on the same probabilities (bit-exact)
    ref_cs = V.segments_from_probs(probs, dict(min_silence_duration_ms=100))
    ref_mask, ref_segs = V.get_segments(ref_cs, pcm)
    assert mask == ref_mask
    assert [(d["start"], d["end"], len(d["samples"])) for d in segs] == [(a, b, len(c)) for a, b, c in ref_segs]
    vad.close()
