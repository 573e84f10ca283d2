"""GPU suite for the diarize flow around the boundary (reference src/engine.rs:117-122, src/transcribe.rs:461-497) as mirrored by
hostmirror.host.diarize: get_segments -> embeddings -> speaker ids.

Bit-exact checks: the labels are the oracle's leader scan (strict >, cap -> best match) / agglomerative clustering applied to the
library's own similarity matrix; short segments (< one fbank frame) get "?".  Floating point: the similarity matrix agrees with the
oracle's cosine of the oracle's embeddings to 5e-3 (bf16 activations in the ResNet)."""
import numpy as np
import pytest

from conftest import synth_audio

pytestmark = pytest.mark.gpu


def test_diarize_flow(wdr, oracle):
    from oracle import cluster as K, resnet
    from hostmirror import host as H
    pcm = synth_audio(61, 35.0, n_speakers=3)
    seg = wdr.Segmenter(seed=1234)
    ex = wdr.EmbeddingExtractor(seed=1234)
    segs = seg.get_segments(pcm)
    assert len(segs) >= 1
    for mode, cap in (("leader", wdr.SIZE_MAX), ("leader", 2), ("agglomerative", wdr.SIZE_MAX)):
        out = H.diarize(seg, ex, pcm, 0.5, cap, mode)
        assert [(o["start"], o["end"]) for o in out] == [(s["start"], s["end"]) for s in segs]
        lens = np.array([len(s["samples"]) for s in segs])
        off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        E, status = ex.compute_batch(np.concatenate([s["samples"] for s in segs]).astype(np.int16), off)
        ok = np.flatnonzero(status == 0)
        assert list(status) == [0 if n >= 400 else -6 for n in lens]
        S = wdr.cosine_matrix(E[ok])
        ref = K.leader_labels(S, 0.5, cap if cap != wdr.SIZE_MAX else 10**9) if mode == "leader" else K.agglomerative_labels(S, 0.5)
        want = ["?"] * len(segs)
        for i, l in zip(ok, ref):
            want[i] = str(int(l)) if l > 0 else "?"
        assert [o["speaker"] for o in out] == want
        if cap == 2:
            assert max(int(x) for x in want if x != "?") <= 2
    # similarity matrix against the all-CPU restatement on a few segments
    rw = resnet.resnet_weights(1234)
    pick = [i for i in ok[:4]]
    ref_e = np.stack([resnet.compute(segs[i]["samples"], rw, oracle.kaldi_fbank) for i in pick])
    S_ref = K.cosine_matrix(ref_e)
    S_got = wdr.cosine_matrix(E[pick])
    assert np.abs(S_ref - S_got).max() < 5e-3
    seg.close()
    ex.close()
