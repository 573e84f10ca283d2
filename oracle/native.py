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


def set_threads(n):
    """OpenMP threads of the C port (overrides OMP_NUM_THREADS); returns the team size a parallel region really gets."""
    L = lib()
    L.oracle_set_threads(int(n))
    return int(L.oracle_threads_in_parallel())


def threads():
    return int(lib().oracle_threads_in_parallel())


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


# ---------------------------------------------------------------------------------------------------
# decoder half (oracle/wdr_oracle_full.c)
# ---------------------------------------------------------------------------------------------------
class TokenData(C.Structure):
    """== whisper_token_data"""
    _fields_ = [("id", C.c_int32), ("tid", C.c_int32), ("p", C.c_float), ("plog", C.c_float), ("pt", C.c_float), ("ptsum", C.c_float),
                ("t0", C.c_int64), ("t1", C.c_int64), ("t_dtw", C.c_int64), ("vlen", C.c_float)]

    def as_tuple(self):
        return (self.id, self.tid, self.p, self.plog, self.pt, self.ptsum, self.t0, self.t1, self.t_dtw, self.vlen)


class Vocab(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("n_vocab", "eot", "sot", "translate", "transcribe", "solm", "prev", "nosp", "not_", "beg",
                                       "lang0", "n_langs", "space")]


class LogitParams(C.Structure):
    _fields_ = [("suppress_blank", C.c_int), ("no_timestamps", C.c_int), ("suppress_nst", C.c_int), ("max_initial_ts", C.c_float)]


def make_vocab(n_vocab):
    from . import vocab as V
    s = V.special_ids(n_vocab)
    return Vocab(**{k: s[k] for k, _ in Vocab._fields_})


_full_ready = False


def _full():
    global _full_ready
    L = lib()
    if not _full_ready:
        f32p, i32p, i64p = C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_int64)
        L.oracle_dec_create.restype = C.c_void_p
        L.oracle_dec_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, f32p, C.c_int]
        L.oracle_dec_free.argtypes = [C.c_void_p]
        L.oracle_dec_set_audio.argtypes = [C.c_void_p, f32p]
        L.oracle_dec_get_self_kv.argtypes = [C.c_void_p, f32p, f32p]
        L.oracle_dec_get_self_kv.restype = None
        L.oracle_dec_set_self_kv.argtypes = [C.c_void_p, f32p, f32p]
        L.oracle_dec_set_self_kv.restype = None
        L.oracle_dec_step.argtypes = [C.c_void_p, C.c_int, C.c_int, f32p, i32p, C.c_int, f32p]
        L.oracle_process_logits.argtypes = [f32p, C.POINTER(Vocab), C.POINTER(LogitParams), C.POINTER(TokenData), C.c_int, C.c_int, C.c_int,
                                            f32p, f32p]
        L.oracle_process_logits.restype = None
        L.oracle_sample_greedy.argtypes = [f32p, f32p, C.POINTER(Vocab)]
        L.oracle_sample_greedy.restype = TokenData
        L.oracle_no_speech_prob.argtypes = [f32p, C.POINTER(Vocab), f32p]
        L.oracle_no_speech_prob.restype = C.c_float
        L.oracle_decode_window.argtypes = [C.c_void_p, C.POINTER(Vocab), C.POINTER(LogitParams), i32p, C.c_int, C.c_int, C.c_int, C.c_int,
                                           C.c_int, C.POINTER(TokenData), i32p, f32p, f32p]
        L.oracle_token_timestamps.argtypes = [C.POINTER(TokenData), C.c_int, C.c_int64, C.c_int64, f32p, f32p, C.c_int, C.c_int, C.c_int,
                                              C.c_float, C.c_float, i64p]
        L.oracle_token_timestamps.restype = None
        L.oracle_dtw_attention.argtypes = [C.c_void_p, i32p, C.c_int, i32p, C.c_int, C.c_int, f32p]
        L.oracle_dtw_stamp.argtypes = [C.POINTER(TokenData), C.c_int, C.c_int, i32p, i32p, C.c_int, C.c_int]
        L.oracle_dtw_stamp.restype = None
        _full_ready = True
    return L


class Decoder:
    """KV-cached Whisper decoder for ONE window.  bf16=True applies libwdr_b200's storage precision."""

    def __init__(self, arch, packed_weights, bf16=False):
        from . import weights as W
        self.a = W.ARCHS[arch]
        self.arch = arch
        self.w = np.ascontiguousarray(packed_weights, np.float32)  # keep alive: the C side borrows it
        self.h = _full().oracle_dec_create(self.a["d"], self.a["n_head"], self.a["n_dec"], self.a["n_vocab"],
                                           self.w.ctypes.data_as(C.POINTER(C.c_float)), int(bf16))
        self.vocab = make_vocab(self.a["n_vocab"])

    def close(self):
        if self.h:
            _full().oracle_dec_free(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def get_kv(self):
        n = self.a["n_dec"] * 448 * self.a["d"]
        k, v = np.empty(n, np.float32), np.empty(n, np.float32)
        _full().oracle_dec_get_self_kv(self.h, k.ctypes.data_as(C.POINTER(C.c_float)), v.ctypes.data_as(C.POINTER(C.c_float)))
        return k, v

    def set_kv(self, kv):
        k, v = kv
        _full().oracle_dec_set_self_kv(self.h, k.ctypes.data_as(C.POINTER(C.c_float)), v.ctypes.data_as(C.POINTER(C.c_float)))

    def set_audio(self, enc):
        e, ep = _f32(enc)
        assert e.shape == (1500, self.a["d"])
        _full().oracle_dec_set_audio(self.h, ep)

    def step(self, token, pos, aheads=None, want_logits=True):
        logits = np.empty(self.a["n_vocab"], np.float32) if want_logits else None
        lp = logits.ctypes.data_as(C.POINTER(C.c_float)) if want_logits else C.POINTER(C.c_float)()
        if aheads:
            ah = np.ascontiguousarray(np.array(aheads, np.int32).reshape(-1))
            pr = np.empty((len(aheads), 1500), np.float32)
            rc = _full().oracle_dec_step(self.h, int(token), int(pos), lp, ah.ctypes.data_as(C.POINTER(C.c_int32)), len(aheads),
                                         pr.ctypes.data_as(C.POINTER(C.c_float)))
            assert rc == 0
            return logits, pr
        rc = _full().oracle_dec_step(self.h, int(token), int(pos), lp, C.POINTER(C.c_int32)(), 0, C.POINTER(C.c_float)())
        assert rc == 0
        return logits

    def decode_window(self, prompt, seek=0, seek_end=2999, single_segment=True, delta_min=10, suppress_blank=True, max_initial_ts=1.0):
        """One seek iteration of whisper_full (greedy, T=0).  Returns dict(tokens=[TokenData...], seek_delta, failed, completed,
        n_sampled, result_len, no_speech_prob, margins)."""
        pr = np.ascontiguousarray(np.array(prompt, np.int32))
        toks = (TokenData * 224)()
        info = np.zeros(8, np.int32)
        nsp = C.c_float(0)
        margins = np.zeros(224, np.float32)
        lp = LogitParams(int(suppress_blank), 0, 0, float(max_initial_ts))
        n = _full().oracle_decode_window(self.h, C.byref(self.vocab), C.byref(lp), pr.ctypes.data_as(C.POINTER(C.c_int32)), len(pr),
                                         int(seek), int(seek_end), int(single_segment), int(delta_min), toks,
                                         info.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(nsp), margins.ctypes.data_as(C.POINTER(C.c_float)))
        return dict(tokens=[toks[i] for i in range(n)], seek_delta=int(info[0]), failed=bool(info[1]), completed=bool(info[2]),
                    n_sampled=int(info[3]), result_len=int(info[5]), no_speech_prob=float(nsp.value), margins=margins[: int(info[3])].copy())

    def dtw_attention(self, seq, aheads, n_audio):
        sq = np.ascontiguousarray(np.array(seq, np.int32))
        ah = np.ascontiguousarray(np.array(aheads, np.int32).reshape(-1))
        out = np.empty((len(aheads), len(sq), n_audio), np.float32)
        rc = _full().oracle_dtw_attention(self.h, sq.ctypes.data_as(C.POINTER(C.c_int32)), len(sq), ah.ctypes.data_as(C.POINTER(C.c_int32)),
                                          len(aheads), int(n_audio), out.ctypes.data_as(C.POINTER(C.c_float)))
        assert rc == 0
        return out


def process_logits(logits, n_vocab, tokens_cur, has_ts, seek_delta, suppress_blank=True, max_initial_ts=1.0):
    """whisper_process_logits + greedy whisper_sample_token on one logits row. tokens_cur: list of token ids sampled so far."""
    v = make_vocab(n_vocab)
    lg = np.ascontiguousarray(logits, np.float32).copy()
    lpb = np.empty(n_vocab, np.float32)
    pb = np.empty(n_vocab, np.float32)
    cur = (TokenData * max(1, len(tokens_cur)))()
    for i, t in enumerate(tokens_cur):
        cur[i].id = int(t)
    lp = LogitParams(int(suppress_blank), 0, 0, float(max_initial_ts))
    f32p = C.POINTER(C.c_float)
    _full().oracle_process_logits(lg.ctypes.data_as(f32p), C.byref(v), C.byref(lp), cur, len(tokens_cur), int(has_ts), int(seek_delta),
                                  lpb.ctypes.data_as(f32p), pb.ctypes.data_as(f32p))
    tok = _full().oracle_sample_greedy(pb.ctypes.data_as(f32p), lpb.ctypes.data_as(f32p), C.byref(v))
    return tok, lg, lpb, pb


def sample_stats(probs, logprobs, n_vocab):
    """tid / pt / ptsum (and the greedy pick) of a processed distribution: oracle_sample_greedy."""
    v = make_vocab(n_vocab)
    f32p = C.POINTER(C.c_float)
    pb = np.ascontiguousarray(probs, np.float32)
    lpb = np.ascontiguousarray(logprobs, np.float32)
    return _full().oracle_sample_greedy(pb.ctypes.data_as(f32p), lpb.ctypes.data_as(f32p), C.byref(v))


def token_timestamps(tokens, t0, t1, vlen, energy, token_beg, token_eot, state3, thold_pt=0.01, thold_ptsum=0.01):
    """whisper_exp_compute_token_level_timestamps on a list of TokenData (modified in place). state3: int64[3] carried."""
    n = len(tokens)
    arr = (TokenData * max(1, n))(*tokens)
    vl = np.ascontiguousarray(vlen, np.float32)
    en = np.ascontiguousarray(energy, np.float32)
    _full().oracle_token_timestamps(arr, n, int(t0), int(t1), vl.ctypes.data_as(C.POINTER(C.c_float)), en.ctypes.data_as(C.POINTER(C.c_float)),
                                    len(en), int(token_beg), int(token_eot), float(thold_pt), float(thold_ptsum),
                                    state3.ctypes.data_as(C.POINTER(C.c_int64)))
    return [arr[i] for i in range(n)]


def dtw_stamp(tokens, token_eot, text_idx, time_idx, seek):
    n = len(tokens)
    arr = (TokenData * max(1, n))(*tokens)
    ti = np.ascontiguousarray(text_idx, np.int32)
    tj = np.ascontiguousarray(time_idx, np.int32)
    _full().oracle_dtw_stamp(arr, n, int(token_eot), ti.ctypes.data_as(C.POINTER(C.c_int32)), tj.ctypes.data_as(C.POINTER(C.c_int32)), len(ti),
                             int(seek))
    return [arr[i] for i in range(n)]
