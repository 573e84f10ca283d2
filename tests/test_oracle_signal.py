"""CPU suite: pins the oracle (oracle/wdr_oracle.c) for the signal-processing stages.

The reference's own tests hold nothing for this path (parity unpinned), so the oracle is
checked against the independent OpenAI-lineage implementations in `transformers` / `torchaudio` where the
semantics coincide, and against hand-derived known answers.
"""
import numpy as np
import pytest
import torch

from conftest import synth_audio


def test_filterbank_matches_transformers(filters80, filters128):
    from transformers.audio_utils import mel_filter_bank
    for f, n in ((filters80, 80), (filters128, 128)):
        ref = mel_filter_bank(201, n, 0.0, 8000.0, 16000, norm="slaney", mel_scale="slaney").T
        assert f.shape == (n, 201)
        assert np.abs(f - ref).max() < 1e-7


@pytest.mark.parametrize("n_mel", [80, 128])
def test_oracle_mel_matches_whisper_feature_extractor(oracle, filters80, filters128, n_mel):
    from transformers import WhisperFeatureExtractor
    filt = filters80 if n_mel == 80 else filters128
    x = synth_audio(7, 6.0).astype(np.float32) / 32768.0
    mel = oracle.log_mel(x, filt)
    assert mel.shape == (n_mel, (len(x) + 480000) // 160)
    fe = WhisperFeatureExtractor(feature_size=n_mel)
    ref = fe(x, sampling_rate=16000, return_tensors="np")["input_features"][0]
    # identical framing for the first 3000 frames (frame 0 starts 200 samples before x[0] in both)
    assert np.abs(mel[:, :3000] - ref).max() < 2e-5


def test_oracle_mel_silence_and_tone(oracle, filters80):
    mel = oracle.log_mel(np.zeros(16000, np.float32), filters80)
    assert np.all(mel == np.float32(-1.5))  # log10(1e-10) = -10 -> (max(-10, -18) + 4) / 4
    t = np.arange(32000) / 16000.0
    tone = (0.5 * np.sin(2 * np.pi * 1000.0 * t)).astype(np.float32)
    raw = oracle.log_mel(tone, filters80, normalize=False)
    peak = raw[:, 50].argmax()
    centre = (filters80[peak] * np.arange(201) * 40.0).sum() / filters80[peak].sum()
    assert abs(centre - 1000.0) < 60.0
    # an impulse has a flat spectrum: every frame that contains it gives hann(n)^2 * sum(filter)
    imp = np.zeros(4000, np.float32)
    imp[1000] = 1.0
    raw = oracle.log_mel(imp, filters80, normalize=False)
    f = 6  # frame 6 covers samples [760, 1160): n = 240
    w = 0.5 * (1 - np.cos(2 * np.pi * 240 / 400))
    expect = np.log10(np.maximum(w * w * filters80.sum(1), 1e-10))
    assert np.abs(raw[:, f] - expect).max() < 1e-4


def test_oracle_median_matches_transformers(oracle):
    from transformers.models.whisper.generation_whisper import _median_filter
    rng = np.random.default_rng(3)
    w = rng.standard_normal((3, 17, 120)).astype(np.float32)
    ref = _median_filter(torch.from_numpy(w), 7).numpy()
    assert np.array_equal(oracle.median_filter(w, 7), ref)
    # ramps and impulses at the borders (reflect indexing: idx<0 -> -idx)
    ramp = np.arange(10, dtype=np.float32)[None, None]
    out = oracle.median_filter(ramp, 7)[0, 0]
    assert out.tolist() == [2.0, 2.0, 2.0, 3.0, 4.0, 5.0, 6.0, 7.0, 7.0, 7.0]
    with pytest.raises(ValueError):
        oracle.median_filter(np.zeros((1, 1, 3), np.float32), 7)


def test_oracle_dtw_matches_transformers(oracle):
    from transformers.models.whisper.generation_whisper import _dynamic_time_warping
    rng = np.random.default_rng(5)
    for (n, m) in ((1, 1), (3, 4), (5, 7), (40, 300), (7, 3)):
        x = rng.standard_normal((n, m)).astype(np.float32)
        ti, tj = oracle.dtw(x)
        rti, rtj = _dynamic_time_warping(x.astype(np.float64))
        assert np.array_equal(ti, rti) and np.array_equal(tj, rtj)
        assert ti[0] == 0 and tj[0] == 0 and ti[-1] == n - 1 and tj[-1] == m - 1
        assert np.all(np.diff(ti) >= 0) and np.all(np.diff(tj) >= 0)


def test_oracle_dtw_known_answers(oracle):
    # all-equal costs: every comparison ties -> trace code 2 (move left) until the column border, then up
    x = np.zeros((3, 4), np.float32)
    ti, tj, cost, trace = oracle.dtw(x, want_matrices=True)
    assert trace[1:, 1:].tolist() == [[0, 2, 2, 2], [1, 2, 2, 2], [1, 2, 2, 2]]
    assert ti.tolist() == [0, 1, 2, 2, 2, 2] and tj.tolist() == [0, 0, 0, 1, 2, 3]
    # a diagonal valley is followed exactly
    x = np.ones((4, 4), np.float32)
    x[np.arange(4), np.arange(4)] = -1.0
    ti, tj = oracle.dtw(x)
    assert ti.tolist() == [0, 1, 2, 3] and tj.tolist() == [0, 1, 2, 3]
    # +inf cells are never stepped on when a finite route exists
    x = np.zeros((2, 3), np.float32)
    x[0, 1] = np.inf
    ti, tj, cost, trace = oracle.dtw(x, want_matrices=True)
    assert np.isfinite(cost[2, 3]) and (0, 1) not in set(zip(ti.tolist(), tj.tolist()))
    assert cost[0, 0] == 0 and np.all(np.isinf(cost[0, 1:])) and np.all(np.isinf(cost[1:, 0]))


def test_oracle_dtw_cost_pipeline(oracle):
    rng = np.random.default_rng(11)
    H, T, A = 4, 12, 50
    w = rng.random((H, T, A)).astype(np.float32)
    x = oracle.dtw_cost(w, sot_len=2, width=7)
    assert x.shape == (T - 3, A)
    mean = w.astype(np.float64).mean(1, keepdims=True)
    std = np.sqrt(((w - mean) ** 2).mean(1, keepdims=True) + 1e-9)
    nrm = ((w - mean) / std).astype(np.float32)
    med = oracle.median_filter(nrm, 7)
    ref = -med.mean(0)[2:-1]
    assert np.abs(x - ref).max() < 1e-5


def test_oracle_fbank_matches_torchaudio(oracle):
    import torchaudio.compliance.kaldi as k
    x = synth_audio(21, 2.5)
    ref = k.fbank(torch.from_numpy(x.astype(np.float32))[None], num_mel_bins=80, frame_length=25, frame_shift=10, dither=0.0,
                  energy_floor=0.0, sample_frequency=16000, window_type="povey", preemphasis_coefficient=0.97,
                  remove_dc_offset=True, snip_edges=True, low_freq=20, high_freq=0, use_log_fbank=True, use_power=True).numpy()
    o = oracle.kaldi_fbank(x.astype(np.float32), 80, subtract_mean=False)
    assert o.shape == ref.shape == (1 + (len(x) - 400) // 160, 80)
    assert np.abs(o - ref).max() < 2e-3
    oc = oracle.kaldi_fbank(x.astype(np.float32), 80, subtract_mean=True)
    assert np.abs(oc - (ref - ref.mean(0, keepdims=True))).max() < 2e-3
    assert oracle.kaldi_fbank(np.zeros(399, np.float32)).shape == (0, 80)


def test_oracle_signal_energy(oracle):
    x = np.array([1, -2, 3, -4, 5], np.float32)
    e = oracle.signal_energy(x, hw=1)
    assert np.allclose(e, np.array([3, 6, 9, 12, 9], np.float32) / 3)


@pytest.mark.parametrize("rate,channels", [(48000, 1), (44100, 1), (8000, 1), (22050, 2), (16000, 1), (32000, 1)])
def test_oracle_resampler_matches_scipy(rate, channels):
    """The polyphase restatement (oracle/resample.py) against scipy.signal.resample_poly with the same Kaiser(5) design."""
    from scipy.signal import resample_poly
    from oracle import resample as R
    rng = np.random.default_rng(rate)
    n = 2 * rate // 5 + 7
    t = np.arange(n) / rate
    mono = 9000 * np.sin(2 * np.pi * 440 * t) + 4000 * np.sin(2 * np.pi * 3100 * t) + rng.normal(0, 300, n)
    x = np.clip(np.rint(np.stack([mono + 50 * c for c in range(channels)], axis=1)), -32768, 32767).astype(np.int16)
    got16, got = R.resample_to_16k(x.reshape(-1), rate, channels)
    up, down = R.ratio(rate)
    xm = x.astype(np.float32).sum(axis=1, dtype=np.float32) * np.float32(1.0 / channels) if channels > 1 else x[:, 0].astype(np.float32)
    ref = resample_poly(xm.astype(np.float64), up, down, window=("kaiser", 5.0))
    assert len(got) == len(ref) == R.n_out(n, rate)
    assert np.abs(got * 32768.0 - ref).max() < 2e-3 * np.abs(ref).max() + 1e-6  # fp32-rounded taps vs scipy's float64 taps
    if rate == 16000:
        assert np.array_equal(got16, x[:, 0])  # identity: one tap of weight 1
