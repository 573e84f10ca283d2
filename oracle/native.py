"""ctypes binding of oracle/wdr_oracle.c (built by oracle/Makefile into oracle/_build)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "wdr_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        f32p, i32p = C.POINTER(C.c_float), C.POINTER(C.c_int32)
        _lib.oracle_mel_n_len.argtypes = [C.c_int]
        _lib.oracle_log_mel.argtypes = [f32p, C.c_int, f32p, C.c_int, C.c_int, f32p]
        _lib.oracle_median_filter.argtypes = [f32p, C.c_int, C.c_int, C.c_int, C.c_int, f32p]
        _lib.oracle_dtw.argtypes = [f32p, C.c_int, C.c_int, i32p, i32p, C.POINTER(C.c_int), f32p, i32p]
        _lib.oracle_dtw_cost.argtypes = [f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, f32p]
        _lib.oracle_fbank_frames.argtypes = [C.c_int]
        _lib.oracle_kaldi_fbank.argtypes = [f32p, C.c_int, C.c_int, C.c_int, f32p]
        _lib.oracle_signal_energy.argtypes = [f32p, C.c_int, C.c_int, f32p]
        _lib.oracle_signal_energy.restype = None
        _lib.oracle_whisper_encode.argtypes = [f32p, C.c_int, C.c_int, C.c_int, C.c_int, f32p, f32p]
    return _lib


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(C.POINTER(C.c_float))


def log_mel(pcm_f32, filters, normalize=True):
    """whisper.cpp log_mel_spectrogram. Returns mel[n_mel, n_len] fp32."""
    pcm, pp = _f32(pcm_f32)
    fl, fp = _f32(filters)
    n_mel = fl.shape[0]
    assert fl.shape[1] == 201
    n_len = lib().oracle_mel_n_len(len(pcm))
    out = np.empty((n_mel, n_len), dtype=np.float32)
    r = lib().oracle_log_mel(pp, len(pcm), fp, n_mel, int(normalize), out.ctypes.data_as(C.POINTER(C.c_float)))
    assert r == n_len
    return out


def median_filter(w, width=7):
    """whisper.cpp median_filter over the last axis of w[H, N, M]."""
    w3, wp = _f32(w)
    H, N, M = w3.shape
    out = np.empty_like(w3)
    rc = lib().oracle_median_filter(wp, H, N, M, width, out.ctypes.data_as(C.POINTER(C.c_float)))
    if rc != 0:
        raise ValueError(f"median_filter rc={rc}")
    return out


def dtw(x, want_matrices=False):
    """whisper.cpp dtw_and_backtrace. Returns (text_idx, time_idx[, cost, trace])."""
    x2, xp = _f32(x)
    N, M = x2.shape
    ti = np.empty(N + M + 2, dtype=np.int32)
    tj = np.empty(N + M + 2, dtype=np.int32)
    n = C.c_int(0)
    cost = trace = None
    cp = C.POINTER(C.c_float)()
    tp = C.POINTER(C.c_int32)()
    if want_matrices:
        cost = np.empty((N + 1, M + 1), dtype=np.float32)
        trace = np.empty((N + 1, M + 1), dtype=np.int32)
        cp = cost.ctypes.data_as(C.POINTER(C.c_float))
        tp = trace.ctypes.data_as(C.POINTER(C.c_int32))
    rc = lib().oracle_dtw(xp, N, M, ti.ctypes.data_as(C.POINTER(C.c_int32)),
                          tj.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(n), cp, tp)
    if rc != 0:
        raise RuntimeError("dtw backtrace hit an unset trace cell")
    if want_matrices:
        return ti[: n.value].copy(), tj[: n.value].copy(), cost, trace
    return ti[: n.value].copy(), tj[: n.value].copy()


def dtw_cost(w, sot_len, width=7):
    """w[H, n_tokens, n_audio] -> cost x[n_tokens-sot_len-1, n_audio]."""
    w3, wp = _f32(w)
    H, T, A = w3.shape
    out = np.empty((T - sot_len - 1, A), dtype=np.float32)
    rc = lib().oracle_dtw_cost(wp, H, T, A, sot_len, width, out.ctypes.data_as(C.POINTER(C.c_float)))
    if rc != 0:
        raise ValueError(f"dtw_cost rc={rc}")
    return out


def kaldi_fbank(wave_i16_scale, n_bins=80, subtract_mean=True):
    """kaldi-native-fbank as pyannote-rs configures it. Returns [T, n_bins] fp32."""
    w, wp = _f32(wave_i16_scale)
    T = lib().oracle_fbank_frames(len(w))
    out = np.empty((T, n_bins), dtype=np.float32)
    if T:
        lib().oracle_kaldi_fbank(wp, len(w), n_bins, int(subtract_mean), out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def signal_energy(signal, hw=32):
    s, sp = _f32(signal)
    out = np.empty_like(s)
    lib().oracle_signal_energy(sp, len(s), hw, out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def whisper_encode(mel, arch, packed_weights):
    """whisper.cpp encoder on one window: mel[n_mel, 3000] normalised -> hidden[1500, d]."""
    from . import weights as W
    a = W.ARCHS[arch]
    m, mp = _f32(mel)
    assert m.shape == (a["n_mel"], 3000)
    pw, pp = _f32(packed_weights)
    out = np.empty((1500, a["d"]), np.float32)
    rc = lib().oracle_whisper_encode(mp, a["n_mel"], a["d"], a["n_head"], a["n_enc"], pp, out.ctypes.data_as(C.POINTER(C.c_float)))
    assert rc == 0
    return out
