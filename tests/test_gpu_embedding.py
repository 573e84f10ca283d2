"""GPU parity of the speaker-embedding path (reference src/transcribe.rs:343, 466-467) through the C ABI:
Kaldi fbank -> WeSpeaker ResNet34 (tcgen05 GEMMs — implicit GEMMs over zero-padded activation maps —, bf16 activations) -> TSTP -> Linear, against oracle/resnet.py
(fp32 activations, the same bf16-rounded folded weights).

Tolerance: the library keeps activations in bf16 between the 36 convolutions, the oracle in fp32, so the embedding agrees to bf16
accumulation error: cosine similarity >= 0.9995 and max |diff| <= 3e-2 of max |embedding| (north-star class for bf16 paths: 1e-2
relative on hidden states; this network is 34 layers deep)."""
import numpy as np
import pytest

from conftest import synth_audio

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rw():
    from oracle import resnet
    return resnet.resnet_weights(1234)


def _check(e, r):
    cos = float(e @ r / (np.linalg.norm(e) * np.linalg.norm(r)))
    rel = float(np.abs(e - r).max() / np.abs(r).max())
    assert cos >= 0.9995 and rel <= 3e-2, (cos, rel)
    return cos, rel


def test_single_segment_matches_oracle(wdr, oracle, rw):
    from oracle import resnet
    ex = wdr.EmbeddingExtractor(seed=1234)
    assert ex.dim == 256
    for seed, secs in ((7, 2.0), (8, 3.37), (9, 0.031)):  # 0.031 s = 496 samples -> a single fbank frame
        pcm = synth_audio(seed, max(secs, 0.1))[: int(secs * 16000)]
        e = ex.compute(pcm)
        r = resnet.compute(pcm, rw, oracle.kaldi_fbank)
        _check(e, r)
    with pytest.raises(wdr.WdrError) as err:
        ex.compute(np.zeros(399, np.int16))
    assert err.value.code == -6  # WDR_ERR_TOO_SHORT: the crate maps it to speaker "?"
    ex.close()


def test_batch_ragged_segments(wdr, oracle, rw):
    from oracle import resnet
    rng = np.random.default_rng(3)
    lens = [16000, 300, 48000, 0, 23456, 400, 70000]
    pcm = np.concatenate([synth_audio(20 + i, max(n, 1600) / 16000.0)[:n] for i, n in enumerate(lens)]).astype(np.int16)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    ex = wdr.EmbeddingExtractor(seed=1234)
    emb, status = ex.compute_batch(pcm, off)
    assert list(status) == [0, -6, 0, -6, 0, 0, 0]
    assert not emb[1].any() and not emb[3].any()
    for s, n in enumerate(lens):
        if n >= 400:
            r = resnet.compute(pcm[off[s]: off[s + 1]], rw, oracle.kaldi_fbank)
            _check(emb[s], r)
    # batching does not change a segment's embedding
    single = ex.compute(pcm[off[2]: off[3]])
    assert np.array_equal(single, emb[2])
    assert ex.last_flops() > 0
    ex.close()
