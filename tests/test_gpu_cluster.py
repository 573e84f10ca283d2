"""GPU parity for row a11: the device cosine matrix (tolerance 1e-5 abs) and the device agglomerative clustering, whose labels
must be IDENTICAL to the oracle's given the same similarity matrix (ties included)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_cosine_matrix_and_agglomerative_labels(wdr):
    from oracle import cluster as K
    rng = np.random.default_rng(5)
    for n, d, k in ((37, 256, 4), (200, 512, 6), (1, 16, 1)):
        cent = rng.standard_normal((k, d))
        E = (cent[rng.integers(0, k, n)] + 0.3 * rng.standard_normal((n, d))).astype(np.float32)
        if n > 2:
            E[2] = 0.0  # zero embedding: similarity 0 by definition
        S = wdr.cosine_matrix(E)
        ref = K.cosine_matrix(E)
        assert np.abs(S - ref).max() < 1e-5
        for thr in (0.2, 0.5, 0.8):
            assert np.array_equal(wdr.cluster_agglomerative(S, thr), K.agglomerative_labels(S, thr)), (n, thr)
            assert np.array_equal(wdr.cluster_leader(S, thr), K.leader_labels(S, thr, 10**9))
    # ties: quantised similarities
    Sq = (np.round(K.cosine_matrix(rng.standard_normal((64, 8)).astype(np.float32)) * 4) / 4).astype(np.float32)
    Sq = np.maximum(Sq, Sq.T)
    assert np.array_equal(wdr.cluster_agglomerative(Sq, 0.25), K.agglomerative_labels(Sq, 0.25))
