"""CPU suite for row a9: the oracle's PyanNet building blocks pinned against torch.nn (Conv1d, MaxPool1d, InstanceNorm1d, LSTM),
and the pyannote-rs speech state machine — oracle KATs plus bit-exact agreement of the library's host implementation."""
import numpy as np
import torch


def test_frontend_and_lstm_match_torch():
    from oracle import pyannet as P
    w = P.pyannet_weights(1234)
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(P.WINDOW) * 3000).astype(np.float32)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    with torch.no_grad():
        y = torch.nn.functional.instance_norm(T(x)[None, None], weight=T(w["wav_norm.weight"]), bias=T(w["wav_norm.bias"]), eps=1e-5)
        for i, stride in enumerate((10, 1, 1)):
            b = T(w[f"conv{i}.bias"]) if f"conv{i}.bias" in w else None
            y = torch.nn.functional.conv1d(y, T(w[f"conv{i}.weight"]), b, stride=stride)
            if i == 0:
                y = y.abs()
            y = torch.nn.functional.max_pool1d(y, 3, 3)
            y = torch.nn.functional.leaky_relu(torch.nn.functional.instance_norm(y, weight=T(w[f"norm{i}.weight"]), bias=T(w[f"norm{i}.bias"]), eps=1e-5))
        lstm = torch.nn.LSTM(60, 128, num_layers=4, bidirectional=True, batch_first=True)
        lstm.load_state_dict({k[5:]: T(v) for k, v in w.items() if k.startswith("lstm.")})
        seq, _ = lstm(y.transpose(1, 2))
        z = torch.nn.functional.leaky_relu(seq @ T(w["linear0.weight"]).T + T(w["linear0.bias"]))
        z = torch.nn.functional.leaky_relu(z @ T(w["linear1.weight"]).T + T(w["linear1.bias"]))
        ref = torch.log_softmax(z @ T(w["classifier.weight"]).T + T(w["classifier.bias"]), -1)[0].numpy()
    got = P.pyannet_forward(x, w)
    assert got.shape == ref.shape == (589, 7)
    assert np.abs(got - ref).max() < 2e-4


def _scores_from_classes(cls):
    sc = np.full((len(cls), 7), -5.0, np.float32)
    sc[np.arange(len(cls)), cls] = -0.1
    return sc


def test_state_machine_kat_and_library(wdr):
    from oracle import pyannet as P
    cls = np.zeros(2 * 589, np.int64)
    cls[10:30] = 3        # speech frames 10..29 -> start 721 + 10*270, end 721 + 30*270
    cls[580:600] = 1      # runs across the window boundary: offsets are absolute
    cls[1170:] = 2        # still speaking at the end: never closed, never emitted
    scores = _scores_from_classes(cls).reshape(2, 589, 7)
    n = 2 * P.WINDOW - 7
    ref = P.segments_from_scores(scores, n)
    # upstream's f64 seconds round trip (offset / sr * sr, truncated) may land one sample below the integer offset
    for (_, _, a, b), (ea, eb) in zip(ref, [(721 + 2700, 721 + 8100), (721 + 580 * 270, 721 + 600 * 270)]):
        assert ea - 1 <= a <= ea and eb - 1 <= b <= eb
        assert a == int(ea / 16000.0 * 16000.0) and b == int(eb / 16000.0 * 16000.0)
    assert ref[0][0] == (721 + 2700) / 16000 and ref[0][1] == (721 + 8100) / 16000
    got = wdr.seg_segments_from_scores(scores, n)
    assert [(g["start"], g["end"], g["i0"], g["i1"]) for g in got] == ref
    # sample ranges are clamped to the ORIGINAL length (start to n - 1, end to n): a short input cuts the second segment
    n_short = 721 + 590 * 270
    ref = P.segments_from_scores(scores, n_short)
    got = wdr.seg_segments_from_scores(scores, n_short)
    assert [(g["start"], g["end"], g["i0"], g["i1"]) for g in got] == ref and ref[1][3] == n_short and ref[1][1] * 16000 > n_short
    ref = P.segments_from_scores(scores, 1000)   # everything beyond the input: start clamps to n - 1, end to n
    got = wdr.seg_segments_from_scores(scores, 1000)
    assert [(g["start"], g["end"], g["i0"], g["i1"]) for g in got] == ref and [(a, b) for _, _, a, b in ref] == [(999, 1000), (999, 1000)]
    # ties: argmax takes the FIRST maximum (class 0 wins a 0-vs-k tie => silence)
    tie = np.zeros((1, 589, 7), np.float32)
    assert wdr.seg_segments_from_scores(tie, P.WINDOW) == [] == P.segments_from_scores(tie, P.WINDOW)
    rng = np.random.default_rng(3)
    for _ in range(10):
        sc = rng.standard_normal((3, 589, 7)).astype(np.float32)
        sc[:, :, 0] += rng.uniform(0, 2)
        ref = P.segments_from_scores(sc, 3 * P.WINDOW)
        got = wdr.seg_segments_from_scores(sc, 3 * P.WINDOW)
        assert [(g["start"], g["end"], g["i0"], g["i1"]) for g in got] == ref
