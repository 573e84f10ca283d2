"""CPU oracle (test infrastructure only) for wdr_resample_i16: rational polyphase resampling to 16 kHz mono.

The reference has no resampler (audio::read_wav, reference src/audio.rs:9-20, rejects everything but 16 kHz mono 16-bit), so
there is no reference arithmetic to follow; the definition is the textbook polyphase FIR that scipy.signal.resample_poly
implements, and tests/test_oracle_signal.py pins this restatement against scipy itself:

    up / down = 16000 / rate (reduced), half = 10 max(up, down), fc = 1 / max(up, down)
    g[k] = sinc(fc (k - half)) kaiser_5(k), h = up g / sum(g);   y[m] = sum_j x[j] h[m down + half - j up]
"""
from math import gcd

import numpy as np


def ratio(sample_rate):
    g = gcd(16000, int(sample_rate))
    return 16000 // g, int(sample_rate) // g


def taps(up, down):
    half = 10 * max(up, down)
    fc = 1.0 / max(up, down)
    k = np.arange(2 * half + 1, dtype=np.float64) - half
    g = np.sinc(fc * k) * np.kaiser(2 * half + 1, 5.0)
    return (up * g / g.sum()).astype(np.float32), half  # the library rounds the taps to fp32 once as well


def n_out(n_frames, sample_rate):
    up, down = ratio(sample_rate)
    return (n_frames * up + down - 1) // down


def resample_to_16k(pcm_i16, sample_rate, channels=1):
    """-> (int16 rounded half-to-even and saturated, float64 unrounded / 32768)"""
    x = np.asarray(pcm_i16, np.int16).reshape(-1, channels).astype(np.float32)
    x = (x.sum(axis=1, dtype=np.float32) * np.float32(1.0 / channels) if channels > 1 else x[:, 0]).astype(np.float64)
    up, down = ratio(sample_rate)
    h, half = taps(up, down)
    h = h.astype(np.float64)
    n = n_out(len(x), sample_rate)
    y = np.zeros(n, np.float64)
    m = np.arange(n, dtype=np.int64)
    idx = m * down + half
    j = idx // up
    k = idx - j * up
    while True:
        ok = (k < len(h))
        if not ok.any():
            break
        inside = ok & (j >= 0) & (j < len(x))
        y[inside] += x[j[inside]] * h[k[inside]]
        k = k + up
        j = j - 1
    return np.clip(np.rint(y), -32768, 32767).astype(np.int16), y / 32768.0
