"""GPU suite: the CUDA path (through the C ABI) against the committed golden vectors of tests/golden/ (independent
implementations: transformers / torchaudio / scipy; see tests/golden/make_golden.py).  Bit-exact for integer/index results
(median filter, DTW paths, cluster labels); stated tolerances for floating point (log-mel 1e-4 abs per north-star; encoder hidden
states 1e-2 relative; fbank 2e-3 abs on int16-scale log energies)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(G, name))


def test_mel_golden(wdr, filters80, filters128):
    g = load("mel.npz")
    for n_mel, filt in ((80, filters80), (128, filters128)):
        fe = wdr.MelFrontend(filt)
        mel = fe.log_mel(g["pcm"])
        assert np.abs(mel[:, :160] - g[f"mel{n_mel}"]).max() < 1e-4


def test_median_and_dtw_golden(wdr):
    g = load("median_dtw.npz")
    assert np.array_equal(wdr.median_filter(g["med_in"], 7), g["med_out"])
    for i in range(int(g["n_dtw"])):
        ti, tj = wdr.dtw(g[f"dtw{i}_x"])
        assert np.array_equal(ti, g[f"dtw{i}_ti"]) and np.array_equal(tj, g[f"dtw{i}_tj"]), i


def test_fbank_golden(wdr):
    g = load("fbank.npz")
    o = wdr.kaldi_fbank(g["pcm"], 80, False)
    assert o.shape == g["fbank"].shape and np.abs(o - g["fbank"]).max() < 2e-3


def test_encoder_golden(wdr):
    g = load("encoder_tiny_en.npz")
    ctx = wdr.Context("tiny.en", seed=1234)
    st = ctx.create_state()
    pcm = g["pcm"]
    win = np.zeros((1, 480000), np.int16)
    win[0, : len(pcm)] = pcm
    h = st.encode_chunks(win, np.array([len(pcm)], np.int32))[0]
    scale = np.abs(g["hidden_rows"]).max()
    assert np.abs(h[g["rows"]] - g["hidden_rows"]).max() < 1e-2 * scale
    assert np.abs(np.linalg.norm(h, axis=1) - g["row_norms"]).max() < 1e-2 * g["row_norms"].max()
    st.close()
    ctx.close()


def test_clustering_golden(wdr):
    g = load("clustering.npz")
    S = wdr.cosine_matrix(g["emb"])
    assert np.abs(S - g["S"]).max() < 1e-5
    assert np.array_equal(wdr.cluster_agglomerative(g["S"], 0.5), g["agglomerative_thr05"])
