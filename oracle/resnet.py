"""WeSpeaker ResNet34 speaker-embedding path restated on the CPU — test infrastructure only.

Reference call sites: `EmbeddingExtractor::new(path)` / `compute(&samples)` (reference src/transcribe.rs:343, 466-467): i16 samples
cast to f32 WITHOUT scaling -> Kaldi fbank (80 bins, 25 ms / 10 ms; oracle/wdr_oracle.c `oracle_kaldi_fbank`) -> per-column mean
subtraction -> ONNX model "feats" [1, T, 80] -> "embs" [1, D].  The reference downloads the CAM++ export (src/engine.rs:91); the
north-star names WeSpeaker **ResNet34** (SURVEY §0.4), which is what is restated here (wespeaker/models/resnet.py, published
architecture):

    x [1, 80(F), T] -> conv3x3(1->32) + BN + ReLU
      -> layer1: 3 x BasicBlock(32, stride 1) -> layer2: 4 x BasicBlock(64, stride 2)
      -> layer3: 6 x BasicBlock(128, stride 2) -> layer4: 3 x BasicBlock(256, stride 2)          [256, 10, ceil(T/8)]
      -> TSTP: reshape [2560, T'] (index c*10 + f), concat(mean_t, sqrt(unbiased var_t + 1e-7)) [5120] -> Linear(5120 -> 256)
    BasicBlock(in, planes, stride): relu(bn2(conv3x3(relu(bn1(conv3x3_stride(x))))) + shortcut(x));
                                    shortcut = bn(conv1x1_stride(x)) when stride != 1 or in != planes, else identity.

Inference-mode BatchNorm is folded into the convolution: w' = bf16(w * g / sqrt(var + 1e-5)), b' = beta - mean * g / sqrt(var + 1e-5)
(fp32 arithmetic, the order written in `fold`).  Weights are the seeded synthetic tensors of oracle/weights.py (no checkpoint can
exist offline); conv matrices are rounded to bf16 once, as the library stores them.  Activations stay fp32 here; the library keeps
them in bf16 between layers (tolerance stated in tests/test_gpu_embedding.py).

PARITY UNPINNED: pyannote-rs / ONNX Runtime / the .onnx export are un-vendored and absent offline.  The convolution arithmetic is
torch.nn.functional.conv2d (an independent implementation), so this file pins the *layout and folding conventions*, not a checkpoint.
"""
import numpy as np

from . import weights as W

F32 = np.float32
EMB_DIM = 256
N_BINS = 80
LAYERS = ((32, 3, 1), (64, 4, 2), (128, 6, 2), (256, 3, 2))  # (planes, blocks, stride of the first block)


def conv_specs():
    """[(name, c_in, c_out, ksize, stride)] in execution order (shortcut convs listed after the block's conv2)."""
    specs = [("conv1", 1, 32, 3, 1)]
    c_in = 32
    for li, (planes, blocks, stride) in enumerate(LAYERS, start=1):
        for bi in range(blocks):
            s = stride if bi == 0 else 1
            specs.append((f"layer{li}.{bi}.conv1", c_in, planes, 3, s))
            specs.append((f"layer{li}.{bi}.conv2", planes, planes, 3, 1))
            if s != 1 or c_in != planes:
                specs.append((f"layer{li}.{bi}.shortcut", c_in, planes, 1, s))
            c_in = planes
    return specs


def fold(w, g, beta, mean, var):
    """BN folding, fp32: s = g / sqrt(var + eps); w' = bf16(w * s[c_out]); b' = beta - mean * s."""
    s = (g / np.sqrt(var + F32(1e-5), dtype=F32)).astype(F32)
    wf = W.bf16_round((w * s[:, None, None, None]).astype(F32))
    bf = (beta - (mean * s).astype(F32)).astype(F32)
    return wf, bf


def resnet_weights(seed=1234):
    """Folded weights: {name: (w [c_out, c_in, k, k] fp32 holding bf16 values, b [c_out] fp32)}, plus 'seg_1': (w [256, 5120], b)."""
    out = {}
    for name, ci, co, k, _ in conv_specs():
        fan = ci * k * k
        w = W.synth(seed, f"resnet34.{name}.weight", (co, ci, k, k), 0.0, np.sqrt(6.0 / fan), native_ok=False)  # He-uniform
        g = W.synth(seed, f"resnet34.{name}.bn.weight", (co,), 0.7 if name.endswith("conv2") or name.endswith("shortcut") else 1.0, 0.1, native_ok=False)
        beta = W.synth(seed, f"resnet34.{name}.bn.bias", (co,), 0.0, 0.1, native_ok=False)
        mean = W.synth(seed, f"resnet34.{name}.bn.running_mean", (co,), 0.0, 0.1, native_ok=False)
        var = W.synth(seed, f"resnet34.{name}.bn.running_var", (co,), 1.0, 0.2, native_ok=False)
        out[name] = fold(w, g, beta, mean, var)
    out["seg_1"] = (W.synth(seed, "resnet34.seg_1.weight", (EMB_DIM, 5120), 0.0, 1.0 / np.sqrt(5120.0), native_ok=False),
                    W.synth(seed, "resnet34.seg_1.bias", (EMB_DIM,), 0.0, 0.05, native_ok=False))
    return out


def resnet_forward(feats, w):
    """feats [T, 80] fp32 (mean-subtracted fbank) -> embedding [256] fp32."""
    import torch
    import torch.nn.functional as TF

    torch.set_grad_enabled(False)

    def conv(x, name, stride, k):
        wt, b = w[name]
        return TF.conv2d(x, torch.from_numpy(wt), torch.from_numpy(b), stride=stride, padding=k // 2)

    x = torch.from_numpy(np.ascontiguousarray(feats, F32).T.copy())[None, None]  # [1, 1, F, T]
    x = torch.relu(conv(x, "conv1", 1, 3))
    c_in = 32
    for li, (planes, blocks, stride) in enumerate(LAYERS, start=1):
        for bi in range(blocks):
            s = stride if bi == 0 else 1
            y = torch.relu(conv(x, f"layer{li}.{bi}.conv1", s, 3))
            y = conv(y, f"layer{li}.{bi}.conv2", 1, 3)
            sc = conv(x, f"layer{li}.{bi}.shortcut", s, 1) if (s != 1 or c_in != planes) else x
            x = torch.relu(y + sc)
            c_in = planes
    _, C, Fq, Tq = x.shape
    v = x.reshape(C * Fq, Tq).double()
    mean = v.mean(1)
    var = ((v - mean[:, None]) ** 2).sum(1) / max(Tq - 1, 1)  # unbiased (torch.var default); a single frame gives 0 rather than NaN
    stats = torch.cat([mean, torch.sqrt(var + 1e-7)]).float()
    lw, lb = w["seg_1"]
    return (torch.from_numpy(lw) @ stats + torch.from_numpy(lb)).numpy().astype(F32)


def compute(pcm_i16, w, fbank):
    """EmbeddingExtractor::compute: int16 samples -> embedding, or None when the segment yields no fbank frame (the crate maps
    that error to speaker "?", reference src/transcribe.rs:468-476).  fbank(pcm_f32, n_bins, subtract_mean) is the oracle's
    Kaldi fbank (oracle.native.kaldi_fbank)."""
    x = np.asarray(pcm_i16, np.int16).astype(F32)
    if len(x) < 400:
        return None
    return resnet_forward(fbank(x, N_BINS, True), w)


def flops(T):
    """Algorithmic conv FLOPs (2 * M * N * K) of one segment with T fbank frames."""
    total = 0
    Fq, Tq = N_BINS, T
    for name, ci, co, k, s in conv_specs():
        if name.endswith("shortcut"):
            fo, to = Fq, Tq  # dims already advanced by the block's conv1
        else:
            fo, to = -(-Fq // s), -(-Tq // s)
        total += 2 * fo * to * co * ci * k * k
        if not name.endswith("shortcut"):
            Fq, Tq = fo, to
    return total + 2 * EMB_DIM * 5120
