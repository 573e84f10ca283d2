"""CPU suite: the kernels' own arithmetic (the __host__ __device__ code in csrc/*_core.cuh, compiled with g++)
against numpy — catches FFT/indexing mistakes before any GPU time is spent."""
import ctypes as C

import numpy as np

fp = C.POINTER(C.c_float)


def _ptr(a):
    return a.ctypes.data_as(fp)


def test_dft16_dft25(harness):
    rng = np.random.default_rng(0)
    for n, fn in ((16, harness.harness_dft16), (25, harness.harness_dft25)):
        x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
        io = x.view(np.float32).copy()
        fn(_ptr(io))
        ref = np.fft.fft(x.astype(np.complex128))
        assert np.abs(io.view(np.complex64) - ref).max() < 5e-6 * np.abs(ref).max()


def test_mel_power_tile(harness):
    rng = np.random.default_rng(1)
    tile = rng.standard_normal(5360).astype(np.float32)
    P = np.zeros((201, 32), np.float32)
    harness.harness_mel_power(_ptr(tile), _ptr(P))
    hann = 0.5 * (1 - np.cos(2 * np.pi * np.arange(400) / 400))
    ref = np.stack([np.abs(np.fft.rfft(hann * tile[f * 160:f * 160 + 400].astype(np.float64))) ** 2 for f in range(32)], 1)
    assert (np.abs(P - ref) / ref.max()).max() < 1e-6


def test_fbank_power_tile(harness):
    rng = np.random.default_rng(2)
    tile = np.round(rng.standard_normal(5360) * 3000 + 500).astype(np.float32)
    P = np.zeros((256, 32), np.float32)
    harness.harness_fbank_power(_ptr(tile), _ptr(P))
    win = (0.5 - 0.5 * np.cos(2 * np.pi / 399 * np.arange(400))) ** 0.85
    ref = []
    for f in range(32):
        x = tile[f * 160:f * 160 + 400].astype(np.float64)
        x = x - x.mean()
        y = x.copy()
        y[1:] -= 0.97 * x[:-1]
        y[0] -= 0.97 * x[0]
        ref.append(np.abs(np.fft.rfft(y * win, 512))[:256] ** 2)
    ref = np.stack(ref, 1)
    assert (np.abs(P - ref) / ref.max()).max() < 1e-6
