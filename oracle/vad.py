"""Silero-VAD path restated in numpy — test infrastructure only.

* silero_probs: whisper.cpp's Silero v5 graph (whisper_vad_build_graph): 512-sample frames, reflect pad 64, STFT as conv1d
  (kernel 256, stride 128, 129 real + 129 imaginary channels), magnitude, 4 x (conv1d k3 + ReLU; strides 1,2,2,1;
  129->128->64->64->128), LSTM cell (128) carried across frames, ReLU, conv1x1 128->1, sigmoid.
* segments_from_probs: whisper_vad_segments_from_probs (threshold / neg_threshold hysteresis, min speech, min silence,
  speech pad, samples -> centiseconds).
* get_segments: the crate's own host logic (reference src/vad.rs:33-82): cs -> s mask, sort, merge gaps < 200 ms, slice the
  int16 samples with f32 rounding.
PARITY UNPINNED (whisper.cpp is un-vendored; no Silero model file offline): weights are seeded synthetic tensors with the
documented shapes, the STFT basis is the true windowed DFT basis."""
import numpy as np

from . import weights as W

F = np.float32
N_WINDOW = 512


def vad_weights(seed=1234):
    """dict of fp32 arrays (PyTorch layouts: conv [out, in, k]; LSTM [4*128, 128] in i,f,g,o order)."""
    w = {}
    k = np.arange(256)
    win = 0.5 * (1 - np.cos(2 * np.pi * k / 256))  # periodic Hann
    c = np.arange(129)[:, None]
    ang = 2 * np.pi * c * k[None, :] / 256
    w["stft.basis"] = np.concatenate([np.cos(ang) * win, -np.sin(ang) * win], 0).astype(F)  # [258, 256]
    chans = [(129, 128), (128, 64), (64, 64), (64, 128)]
    for i, (ci, co) in enumerate(chans):
        s = 1.0 / np.sqrt(ci * 3)
        w[f"enc.{i}.weight"] = W.synth(seed, f"vad.encoder.{i}.weight", (co, ci, 3), 0.0, s, native_ok=False)
        w[f"enc.{i}.bias"] = W.synth(seed, f"vad.encoder.{i}.bias", (co,), 0.0, s, native_ok=False)
    s = 1.0 / np.sqrt(128)
    w["lstm.w_ih"] = W.synth(seed, "vad.lstm.weight_ih", (512, 128), 0.0, s, native_ok=False)
    w["lstm.w_hh"] = W.synth(seed, "vad.lstm.weight_hh", (512, 128), 0.0, s, native_ok=False)
    w["lstm.b_ih"] = W.synth(seed, "vad.lstm.bias_ih", (512,), 0.0, s, native_ok=False)
    w["lstm.b_hh"] = W.synth(seed, "vad.lstm.bias_hh", (512,), 0.0, s, native_ok=False)
    # The output head is drawn 250x wider than a default init and re-centred: with default-init scales a random Silero's probability
    # stays inside [0.484, 0.489] whatever the audio (it never crosses the 0.5 / 0.35 hysteresis, so the segmenter would have nothing
    # to do); with these constants it follows the signal's energy across the thresholds (csrc/vad.cu uses the same constants).
    w["final.weight"] = W.synth(seed, "vad.final_conv.weight", (128,), 0.0, 125.0, native_ok=False)
    w["final.bias"] = W.synth(seed, "vad.final_conv.bias", (1,), 14.2, 25.0, native_ok=False)
    return w


def _conv1d(x, w, b, stride, pad):
    """x [C_in, T], w [C_out, C_in, K] -> [C_out, T_out], fp32."""
    ci, T = x.shape
    co, _, K = w.shape
    xp = np.zeros((ci, T + 2 * pad), F)
    xp[:, pad:pad + T] = x
    To = (T + 2 * pad - K) // stride + 1
    out = np.empty((co, To), F)
    for t in range(To):
        out[:, t] = np.tensordot(w, xp[:, t * stride:t * stride + K], axes=([1, 2], [0, 1])) + b
    return out


def frame_features(frame, w):
    """One 512-sample frame -> the 128-vector fed to the LSTM."""
    x = np.asarray(frame, F)
    xp = np.concatenate([x[64:0:-1], x, x[-2:-66:-1]])  # reflect pad 64 | 64 -> 640
    cols = np.stack([xp[t * 128:t * 128 + 256] for t in range(4)], 1)  # [256, 4]
    st = (w["stft.basis"] @ cols).astype(F)  # [258, 4]
    mag = np.sqrt(st[:129] ** 2 + st[129:] ** 2).astype(F)
    cur = mag
    for i, stride in enumerate((1, 2, 2, 1)):
        cur = np.maximum(_conv1d(cur, w[f"enc.{i}.weight"], w[f"enc.{i}.bias"], stride, 1), 0).astype(F)
    return cur[:, 0]


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def silero_probs(pcm_f32, w):
    """whisper_vad_detect_speech: per-frame speech probability; LSTM state reset at the start of the call."""
    x = np.asarray(pcm_f32, F)
    n = len(x)
    n_frames = (n + N_WINDOW - 1) // N_WINDOW
    h = np.zeros(128, F)
    c = np.zeros(128, F)
    probs = np.zeros(n_frames, F)
    for i in range(n_frames):
        fr = np.zeros(N_WINDOW, F)
        seg = x[i * N_WINDOW:(i + 1) * N_WINDOW]
        fr[:len(seg)] = seg
        feat = frame_features(fr, w)
        g = (w["lstm.w_ih"] @ feat + w["lstm.b_ih"] + w["lstm.w_hh"] @ h + w["lstm.b_hh"]).astype(F)
        ig, fg, gg, og = _sigmoid(g[:128]), _sigmoid(g[128:256]), np.tanh(g[256:384]), _sigmoid(g[384:])
        c = (fg * c + ig * gg).astype(F)
        h = (og * np.tanh(c)).astype(F)
        y = F(np.dot(w["final.weight"], np.maximum(h, 0)) + w["final.bias"][0])
        probs[i] = _sigmoid(y)
    return probs


def default_params():
    return dict(threshold=0.5, min_speech_duration_ms=250, min_silence_duration_ms=100, max_speech_duration_s=3.4e38, speech_pad_ms=30,
                samples_overlap=0.1)


def samples_to_cs(samples):
    return int((samples / 16000.0) * 100.0 + 0.5)


def segments_from_probs(probs, p=None):
    """whisper_vad_segments_from_probs -> list of (start_cs, end_cs) floats (the values whisper_vad_segments_get_segment_t0/_t1 return)."""
    p = dict(default_params(), **(p or {}))
    probs = np.asarray(probs, F)
    n_probs = len(probs)
    sr = 16000
    threshold = F(p["threshold"])
    min_silence_samples = sr * p["min_silence_duration_ms"] // 1000
    audio_length_samples = n_probs * N_WINDOW
    min_speech_samples = sr * p["min_speech_duration_ms"] // 1000
    speech_pad_samples = sr * p["speech_pad_ms"] // 1000
    if p["max_speech_duration_s"] > 100000.0:
        max_speech_samples = (2 ** 31 - 1) // 2
    else:
        max_speech_samples = int(sr * int(p["max_speech_duration_s"]) - N_WINDOW - 2 * speech_pad_samples)
        if max_speech_samples < 0:
            max_speech_samples = (2 ** 31 - 1) // 2
    min_silence_at_max = sr * 98 // 1000
    neg_threshold = F(threshold - F(0.15))
    if neg_threshold < F(0.01):
        neg_threshold = F(0.01)
    speeches = []
    is_speech = False
    temp_end = prev_end = next_start = curr_start = 0
    has_curr = False
    for i in range(n_probs):
        pr = probs[i]
        cs = N_WINDOW * i
        if pr >= threshold and temp_end:
            temp_end = 0
            if next_start < prev_end:
                next_start = cs
        if pr >= threshold and not is_speech:
            is_speech = True
            curr_start = cs
            has_curr = True
            continue
        if is_speech and (cs - curr_start) > max_speech_samples:
            if prev_end:
                speeches.append([curr_start, prev_end])
                has_curr = True
                if next_start < prev_end:
                    is_speech = False
                    has_curr = False
                else:
                    curr_start = next_start
                prev_end = next_start = temp_end = 0
            else:
                speeches.append([curr_start, cs])
                prev_end = next_start = temp_end = 0
                is_speech = False
                has_curr = False
                continue
        if pr < neg_threshold and is_speech:
            if not temp_end:
                temp_end = cs
            if (cs - temp_end) > min_silence_at_max:
                prev_end = temp_end
            if (cs - temp_end) < min_silence_samples:
                continue
            if (temp_end - curr_start) > min_speech_samples:
                speeches.append([curr_start, temp_end])
            prev_end = next_start = temp_end = 0
            is_speech = False
            has_curr = False
            continue
    if has_curr and (audio_length_samples - curr_start) > min_speech_samples:
        speeches.append([curr_start, audio_length_samples])
    for i in range(len(speeches)):
        if i == 0:
            speeches[i][0] = speeches[i][0] - speech_pad_samples if speeches[i][0] > speech_pad_samples else 0
        if i < len(speeches) - 1:
            sil = speeches[i + 1][0] - speeches[i][1]
            if sil < 2 * speech_pad_samples:
                speeches[i][1] += sil // 2
                speeches[i + 1][0] = speeches[i + 1][0] - sil // 2 if speeches[i + 1][0] > sil // 2 else 0
            else:
                speeches[i][1] = min(speeches[i][1] + speech_pad_samples, audio_length_samples)
                speeches[i + 1][0] = speeches[i + 1][0] - speech_pad_samples if speeches[i + 1][0] > speech_pad_samples else 0
        else:
            speeches[i][1] = min(speeches[i][1] + speech_pad_samples, audio_length_samples)
    return [(float(F(samples_to_cs(s))), float(F(samples_to_cs(e)))) for s, e in speeches]


def get_segments(segs_cs, int_samples):
    """reference src/vad.rs:33-82 on the VAD's (start_cs, end_cs) list: returns (mask, [(start_s, end_s, samples)])."""
    n = len(int_samples)
    SR = F(16000.0)
    n_f32 = F(n)
    mask = [(float(F(s)) / 100.0, float(F(e)) / 100.0) for s, e in segs_cs]
    mask = [(s, e) for s, e in mask if e > s]
    mask.sort(key=lambda t: t[0])
    merged = []
    for s, e in mask:
        if merged and s - merged[-1][1] < 0.200:
            merged[-1][1] = max(e, merged[-1][1])
        else:
            merged.append([s, e])
    out = []
    for s, e in merged:
        si = int(np.clip(np.floor(F(F(s) * SR) + F(0.5)), F(0.0), n_f32))  # f32::round (half away from zero; values are >= 0)
        ei = int(np.clip(np.floor(F(F(e) * SR) + F(0.5)), F(0.0), n_f32))
        seg = int_samples[si:ei] if ei > si else int_samples[:0]
        if e > s and len(seg):
            out.append((s, e, seg))
    return mask, out
