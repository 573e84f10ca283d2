"""GPU suite at BASELINE.json's full size (configs[2]: large-v3, 128 mel bins, 120 x 30 s windows per GPU), where the CPU oracle
would need ~1 h: size-independent properties of the whole sharded call instead of a token-by-token comparison.

* determinism: two calls give identical segments, tokens, statistics and times;
* batch-composition invariance: a window decoded in a batch of 120 equals the same window decoded alone / in a batch of 3 (every
  kernel's result for a window is independent of the other rows of its launch) — this ties the full-size run to the small-batch
  runs that ARE checked against the oracle token by token (tests/test_gpu_decoder.py);
* structure: chunk order, one segment per window (single_segment), token ids inside the vocabulary, text tokens < EOT counted by
  n_tokens, segment bounds inside the window, 0 <= t0 <= t1 <= 3000 per token, DTW times non-decreasing
  within a segment and inside the window, probabilities in (0, 1], plog = log(p) for unmasked picks."""
import numpy as np
import pytest

from conftest import synth_audio

pytestmark = pytest.mark.gpu


def _key(s):
    return (s["t0"], s["t1"], s["text"], [(t.id, t.tid, t.p, t.plog, t.pt, t.ptsum, t.t0, t.t1, t.t_dtw, t.vlen) for t in s["tokens"]])


def test_large_v3_120_windows_properties(wdr):
    B = 120
    base = [synth_audio(2000 + i, 30.0) for i in range(4)]
    pcm = np.empty((B, 480000), np.int16)
    for i in range(B):
        pcm[i] = np.roll(base[i % 4], 7919 * (i // 4))
    ctx = wdr.Context("large-v3", seed=1234, enable_dtw=True)
    st = ctx.create_state()
    nv_vocab = ctx.dims.n_vocab
    eot, beg = 50257, nv_vocab - 1501
    segs = st.full_batch(pcm)
    again = st.full_batch(pcm)
    assert [_key(s) for s in segs] == [_key(s) for s in again], "the call is not deterministic"
    chunks = [s["chunk"] for s in segs]
    assert chunks == sorted(chunks) and len(set(chunks)) == len(chunks), "one segment per window, in window order"
    assert len(segs) >= B // 2
    n_dtw = 0
    for s in segs:
        info = st.chunk_info(s["chunk"])
        assert 0 <= s["t0"] <= s["t1"] <= 3000 and s["t1"] == info["seek_delta"]
        assert len(s["tokens"]) == (info["n_sampled"] if info["failed"] else info["result_len"])
        last_t1, last_dtw = 0, -1
        for t in s["tokens"]:
            assert 0 <= t.id < nv_vocab and (t.tid == 0 or beg <= t.tid < nv_vocab)  # tid = 0: every timestamp token was masked
            assert 0.0 < t.p <= 1.0 and abs(np.log(t.p) - t.plog) < 1e-3 * max(1.0, abs(t.plog)) and 0.0 <= t.ptsum <= 1.0 + 1e-5
            assert 0 <= t.t0 <= t.t1 <= 3000, (s["chunk"], t.id, t.t0, t.t1)
            if t.id < eot:
                if t.t_dtw >= 0:
                    assert last_dtw <= t.t_dtw <= 3000
                    last_dtw = t.t_dtw
                    n_dtw += 1
            else:
                assert t.t_dtw == -1
        assert s["text"] == "".join(tt for tt, t in zip(s["token_text"], s["tokens"]) if t.id < eot)
    assert n_dtw > 1000
    # batch-composition invariance
    by_chunk = {s["chunk"]: s for s in segs}
    for pick in ([7], [0, 63, 119]):
        sub = st.full_batch(pcm[pick])
        for k, s in enumerate(sub):
            ref = by_chunk.get(pick[s["chunk"]])
            assert ref is not None and _key(s) == _key(ref), pick
        assert len(sub) == sum(1 for c in pick if c in by_chunk)
    st.close()
    ctx.close()
