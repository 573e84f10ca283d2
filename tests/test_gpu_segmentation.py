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
    assert seg.scores(np.zeros(0, np.int16)).shape == (0, 589, 7)
    seg.close()


def test_get_segments_chain(wdr):
    from oracle import pyannet as P
    seg = wdr.Segmenter(seed=1234)
    pcm = synth_audio(52, 12.0, n_speakers=2)
    scores = seg.scores(pcm)
    got = seg.get_segments(pcm)
    ref = P.segments_from_scores(scores, 2 * P.WINDOW)
    assert [(g["start"], g["end"], g["i0"], g["i1"]) for g in got] == ref
    padded = np.zeros(2 * P.WINDOW, np.int16)
    padded[: len(pcm)] = pcm
    for g in got:
        assert np.array_equal(g["samples"], padded[g["i0"]:g["i1"]])
    seg.close()
