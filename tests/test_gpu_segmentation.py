"""GPU parity for row a9: PyanNet window scores on the device against the numpy oracle (fp32 network with wide recurrent weights:
log-probabilities of magnitude ~20 agree to 5e-3 absolute, per-frame argmax on >= 99.5 % of the frames) incl. a ragged last window, and the pyannote_rs::get_segments chain (state machine bit-exact on the device's
own scores)."""
import numpy as np
import pytest

from conftest import synth_audio

pytestmark = pytest.mark.gpu


def test_scores_match_oracle(wdr):
    from oracle import pyannet as P
    w = P.pyannet_weights(1234)
    seg = wdr.Segmenter(seed=1234)
    pcm = synth_audio(51, 23.7, n_speakers=2)  # 3 windows, the last one zero padded
    got = seg.scores(pcm)
    assert got.shape == (3, 589, 7)
    padded = np.zeros(3 * P.WINDOW, np.int16)
    padded[: len(pcm)] = pcm
    for i in range(3):
        ref = P.pyannet_forward(padded[i * P.WINDOW:(i + 1) * P.WINDOW].astype(np.float32), w)
        assert np.abs(got[i] - ref).max() < 5e-3, (i, np.abs(got[i] - ref).max())
        assert (got[i].argmax(-1) == ref.argmax(-1)).mean() >= 0.995
    assert seg.scores(np.zeros(0, np.int16)).shape == (1, 589, 7)  # pyannote-rs pads window - 0 % window zeros: one silent window
    seg.close()


def test_get_segments_chain(wdr):
    from oracle import pyannet as P
    seg = wdr.Segmenter(seed=1234)
    pcm = synth_audio(52, 12.0, n_speakers=2)
    scores = seg.scores(pcm)
    got = seg.get_segments(pcm)
    ref = P.segments_from_scores(scores, len(pcm))
    assert [(g["start"], g["end"], g["i0"], g["i1"]) for g in got] == ref
    for g in got:
        assert g["i1"] <= len(pcm) and np.array_equal(g["samples"], pcm[g["i0"]:g["i1"]])  # clamped to the input: no padding zeros
    seg.close()


def test_exact_multiple_of_window_gets_a_silent_extra_window(wdr):
    """n == k * 160000: pyannote-rs pads a WHOLE extra window (`window - len % window`), whose silence closes a speaker still active at
    the end of the audio; the segment's samples stop at the input's end."""
    from oracle import pyannet as P
    seg = wdr.Segmenter(seed=1234)
    pcm = synth_audio(53, 20.0, n_speakers=2)[: 2 * P.WINDOW]
    assert len(pcm) == 2 * P.WINDOW and wdr.load().wdr_seg_n_windows(len(pcm)) == 3
    scores = seg.scores(pcm)
    assert scores.shape == (3, 589, 7)
    got = seg.get_segments(pcm)
    ref = P.segments_from_scores(scores, len(pcm))
    assert [(g["start"], g["end"], g["i0"], g["i1"]) for g in got] == ref
    assert all(g["i1"] <= len(pcm) for g in got)
    # the state machine on scores that speak through the end of window 1 and fall silent in the extra window
    cls = np.zeros(3 * 589, np.int64)
    cls[1100:1178 + 5] = 2
    sc = np.full((3 * 589, 7), -5.0, np.float32)
    sc[np.arange(len(cls)), cls] = -0.1
    ref = P.segments_from_scores(sc.reshape(3, 589, 7), len(pcm))
    got = wdr.seg_segments_from_scores(sc.reshape(3, 589, 7), len(pcm))
    assert [(g["start"], g["end"], g["i0"], g["i1"]) for g in got] == ref and len(ref) == 1
    assert ref[0][3] == len(pcm) and ref[0][1] * 16000 > len(pcm)  # end TIME runs into the padding, the sample range does not
    seg.close()
