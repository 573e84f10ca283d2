import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import native
    native.build()
    return native


@pytest.fixture(scope="session")
def filters80():
    from oracle import filters
    return filters.whisper_mel_filters(80)


@pytest.fixture(scope="session")
def filters128():
    from oracle import filters
    return filters.whisper_mel_filters(128)


@pytest.fixture(scope="session")
def wdr():
    import wdr_b200
    wdr_b200.load()
    return wdr_b200


@pytest.fixture(scope="session")
def harness():
    """The kernels' __host__ __device__ arithmetic compiled with g++ (tests/host_harness.cpp)."""
    here = os.path.dirname(os.path.abspath(__file__))
    out = os.path.join(here, "_build", "libharness.so")
    src = os.path.join(here, "host_harness.cpp")
    inc = os.path.join(ROOT, "whisper-diarize-rs_b200", "csrc")
    deps = [src] + [os.path.join(inc, f) for f in os.listdir(inc) if f.endswith(".cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-x", "c++", "-Wno-unknown-pragmas", "-I", inc, src, "-o", out])
    return ctypes.CDLL(out)


def synth_audio(seed, seconds, n_speakers=1, sr=16000):
    """Seeded synthetic 'speech' (SURVEY §8d): harmonic sources through formant-ish resonators, syllable AM,
    speaker turns with silences, -40 dBFS noise, peak -3 dBFS.  Returns int16 mono."""
    rng = np.random.default_rng(seed)
    n = int(seconds * sr)
    t = np.arange(n) / sr
    out = np.zeros(n)
    f0s = [110.0, 150.0, 190.0, 230.0]
    pos = 0
    spk = 0
    while pos < n:
        dur = int(rng.uniform(2.0, 8.0) * sr)
        sil = int(rng.uniform(0.3, 1.0) * sr)
        end = min(n, pos + dur)
        tt = t[pos:end]
        f0 = f0s[spk % 4] * (1 + 0.05 * np.sin(2 * np.pi * 5.0 * tt))
        ph = 2 * np.pi * np.cumsum(f0) / sr
        formants = rng.uniform(300, 3000, size=3)
        sig = np.zeros(end - pos)
        for k in range(1, 13):
            fk = f0s[spk % 4] * k
            gain = sum(1.0 / (1.0 + ((fk - fc) / 150.0) ** 2) for fc in formants)
            sig += (gain / k) * np.sin(k * ph)
        am = 0.5 * (1 + np.sin(2 * np.pi * rng.uniform(3, 5) * tt + rng.uniform(0, 6.28)))
        out[pos:end] = sig * am
        pos = end + sil
        spk = (spk + 1) % max(1, n_speakers)
    out += 10 ** (-40 / 20) * rng.standard_normal(n) * (np.abs(out).max() + 1e-9)
    out *= 10 ** (-3 / 20) / (np.abs(out).max() + 1e-9)
    return np.round(out * 32767).astype(np.int16)
