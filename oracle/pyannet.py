"""pyannote segmentation-3.0 path restated in numpy — test infrastructure only.

* pyannet_forward: PyanNet = SincNet front end (InstanceNorm(1) -> conv 1->80 k251 s10 -> |.| -> maxpool3 -> InstanceNorm(80) ->
  LeakyReLU; conv 80->60 k5 -> maxpool3 -> IN -> LeakyReLU; conv 60->60 k5 -> maxpool3 -> IN -> LeakyReLU) -> 4 x biLSTM(128) ->
  Linear 256->128 -> LeakyReLU -> Linear 128->128 -> LeakyReLU -> Linear 128->7 -> log-softmax; 160000 samples -> [589, 7].
* get_segments: pyannote-rs `get_segments`: 10 s windows of RAW int16 values cast to f32, zero-padded to a multiple of the
  window, per-frame argmax != 0 speech state machine (frame_start 721, frame_size 270 samples, absolute across windows).
PARITY UNPINNED (pyannote-rs / ONNX Runtime are un-vendored, no .onnx offline): seeded synthetic weights with the documented
shapes; cross-checked against torch.nn building blocks in tests/test_oracle_pyannet.py."""
import numpy as np

from . import weights as W

F = np.float32
WINDOW = 160000
N_FRAMES = 589
FRAME_START, FRAME_SIZE = 721, 270


def pyannet_weights(seed=1234):
    w = {}

    def t(name, shape, off, sc):
        w[name] = W.synth(seed, "pyannet." + name, shape, off, sc, native_ok=False)

    t("wav_norm.weight", (1,), 1.0, 0.1)
    t("wav_norm.bias", (1,), 0.0, 0.1)
    for i, (ci, co, k) in enumerate(((1, 80, 251), (80, 60, 5), (60, 60, 5))):
        s = 1.0 / np.sqrt(ci * k)
        t(f"conv{i}.weight", (co, ci, k), 0.0, s)
        if i > 0:
            t(f"conv{i}.bias", (co,), 0.0, s)  # the sinc filterbank (conv0) has no bias
        t(f"norm{i}.weight", (co,), 1.0, 0.1)
        t(f"norm{i}.bias", (co,), 0.0, 0.1)
    # Recurrent / head matrices 2.5x / 2x / 6x wider than PyTorch's default init and +1.5 bias on the "no speaker" class: with
    # default-init scales a random PyanNet is constant in time and the state machine never emits a segment (csrc/segmentation.cu
    # uses the same constants).
    s = 1.0 / np.sqrt(128)
    for l in range(4):
        n_in = 60 if l == 0 else 256
        for d in ("", "_reverse"):
            t(f"lstm.weight_ih_l{l}{d}", (512, n_in), 0.0, 2.5 / np.sqrt(128))
            t(f"lstm.weight_hh_l{l}{d}", (512, 128), 0.0, 2.5 / np.sqrt(128))
            t(f"lstm.bias_ih_l{l}{d}", (512,), 0.0, s)
            t(f"lstm.bias_hh_l{l}{d}", (512,), 0.0, s)
    t("linear0.weight", (128, 256), 0.0, 2.0 / 16)
    t("linear0.bias", (128,), 0.0, 1.0 / 16)
    t("linear1.weight", (128, 128), 0.0, 2.0 / np.sqrt(128))
    t("linear1.bias", (128,), 0.0, s)
    t("classifier.weight", (7, 128), 0.0, 24.0 / np.sqrt(128))
    t("classifier.bias", (7,), 0.0, 0.5)
    w["classifier.bias"][0] += np.float32(1.5)
    return w


def _instance_norm(x, g, b, eps=1e-5):
    """x [C, T]: per-channel normalisation over T (biased variance), affine."""
    m = x.mean(1, keepdims=True, dtype=np.float64)
    v = ((x - m) ** 2).mean(1, keepdims=True, dtype=np.float64)
    return (((x - m) / np.sqrt(v + eps)) * g[:, None] + b[:, None]).astype(F)


def _conv1d(x, w, b, stride):
    ci, T = x.shape
    co, _, K = w.shape
    To = (T - K) // stride + 1
    win = np.lib.stride_tricks.as_strided(x, (ci, To, K), (x.strides[0], x.strides[1] * stride, x.strides[1]))
    out = np.einsum("ctk,ock->ot", win, w, optimize=True).astype(F)
    if b is not None:
        out += b[:, None]
    return out


def _maxpool3(x):
    T = x.shape[1] // 3
    return x[:, : T * 3].reshape(x.shape[0], T, 3).max(2)


def _leaky(x):
    return np.where(x > 0, x, F(0.01) * x).astype(F)


def _sigmoid(x):
    return (1.0 / (1.0 + np.exp(-x))).astype(F)


def _lstm_dir(x, w_ih, w_hh, b_ih, b_hh, reverse):
    T = x.shape[0]
    g_in = (x @ w_ih.T + b_ih + b_hh).astype(F)
    h = np.zeros(128, F)
    c = np.zeros(128, F)
    out = np.empty((T, 128), F)
    order = range(T - 1, -1, -1) if reverse else range(T)
    for t in order:
        g = (g_in[t] + w_hh @ h).astype(F)
        i, f, gg, o = _sigmoid(g[:128]), _sigmoid(g[128:256]), np.tanh(g[256:384]).astype(F), _sigmoid(g[384:])
        c = (f * c + i * gg).astype(F)
        h = (o * np.tanh(c)).astype(F)
        out[t] = h
    return out


def pyannet_forward(window_f32, w):
    """window [160000] fp32 (int16-scale values, as pyannote-rs feeds them) -> log-probabilities [589, 7]."""
    x = np.asarray(window_f32, F)[None, :]
    assert x.shape[1] == WINDOW
    x = _instance_norm(x, w["wav_norm.weight"], w["wav_norm.bias"])
    for i, stride in enumerate((10, 1, 1)):
        x = _conv1d(np.ascontiguousarray(x), w[f"conv{i}.weight"], w.get(f"conv{i}.bias"), stride)
        if i == 0:
            x = np.abs(x)
        x = _leaky(_instance_norm(_maxpool3(x), w[f"norm{i}.weight"], w[f"norm{i}.bias"]))
    seq = np.ascontiguousarray(x.T)  # [589, 60]
    assert seq.shape[0] == N_FRAMES
    for l in range(4):
        fw = _lstm_dir(seq, w[f"lstm.weight_ih_l{l}"], w[f"lstm.weight_hh_l{l}"], w[f"lstm.bias_ih_l{l}"], w[f"lstm.bias_hh_l{l}"], False)
        bw = _lstm_dir(seq, w[f"lstm.weight_ih_l{l}_reverse"], w[f"lstm.weight_hh_l{l}_reverse"], w[f"lstm.bias_ih_l{l}_reverse"],
                       w[f"lstm.bias_hh_l{l}_reverse"], True)
        seq = np.concatenate([fw, bw], 1)
    y = _leaky((seq @ w["linear0.weight"].T + w["linear0.bias"]).astype(F))
    y = _leaky((y @ w["linear1.weight"].T + w["linear1.bias"]).astype(F))
    z = (y @ w["classifier.weight"].T + w["classifier.bias"]).astype(F)
    m = z.max(1, keepdims=True)
    return (z - m - np.log(np.exp(z - m).sum(1, keepdims=True))).astype(F)


def segments_from_scores(scores, n_samples, sample_rate=16000):
    """pyannote-rs state machine over consecutive windows' scores [n_windows, 589, 7] -> [(start_s, end_s, start_idx, end_idx)].
    n_samples = the ORIGINAL sample count: upstream computes `start = start_offset / sr`, `start_f64 = start * sr`,
    `start_idx = start_f64.min((len - 1) as f64) as usize`, `end_idx = end_f64.min(len as f64) as usize` (f64 round trip, truncation)."""
    segs = []
    offset = FRAME_START
    speaking = False
    start = 0
    for win in scores:
        for fr in win:
            cls = int(np.argmax(fr))  # first maximum
            if cls != 0:
                if not speaking:
                    start = offset
                    speaking = True
            elif speaking:
                t0, t1 = float(start) / float(sample_rate), float(offset) / float(sample_rate)
                s_idx = int(min(t0 * float(sample_rate), float(max(n_samples - 1, 0))))
                e_idx = max(s_idx, int(min(t1 * float(sample_rate), float(n_samples))))
                segs.append((t0, t1, s_idx, e_idx))
                speaking = False
            offset += FRAME_SIZE
    return segs


def get_segments(int_samples, w):
    """pyannote_rs::get_segments(&samples, 16000, model) (reference src/engine.rs:117-122)."""
    x = np.asarray(int_samples, np.int16)
    n_win = len(x) // WINDOW + 1  # padded.extend(vec![0; window_size - (len % window_size)]): a whole extra window on exact multiples
    padded = np.zeros(n_win * WINDOW, np.int16)
    padded[: len(x)] = x
    scores = np.stack([pyannet_forward(padded[i * WINDOW:(i + 1) * WINDOW].astype(F), w) for i in range(n_win)])
    segs = segments_from_scores(scores, len(x))
    return [dict(start=s, end=e, samples=padded[a:b]) for s, e, a, b in segs], scores
