"""Test-side WRITERS of the model files the crate opens besides ggml-<model>.bin (test infrastructure, like tests/ggml_writer.py):

* `export_pyannet_onnx`  — segmentation-3.0.onnx: a torch module of the PyanNet architecture (SincNet front end with the filterbank as a
  plain Conv1d, as an eval-mode export with constant folding leaves it; 4 x biLSTM(128); 3 Linear; LogSoftmax) carrying the oracle's
  seeded tensors, written by torch's own TorchScript ONNX exporter — the real protobuf torch.onnx.export produces (initializers with
  raw_data, LSTM operators with W / R / B in ONNX gate order i,o,f,c, MatMul + Add for the 3-D Linear layers), not a hand-rolled file.
* `export_resnet34_onnx` — the WeSpeaker ResNet34 export ("feats" [B, T, 80] -> "embs" [B, 256]); `fold_bn` False keeps the
  BatchNorm layers out of eval-mode folding by exporting in training mode so that BatchNormalization nodes stay in the graph.
* `write_silero_ggml`    — ggml-silero-v5.1.2.bin as whisper.cpp's models/convert-silero-vad-to-ggml.py lays it out.

No `onnx` python package exists in this image; the exporter only needs it for onnxscript custom functions, which these modules do
not use, so that one hook is stubbed."""
import io
import struct
import warnings

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as TF


def _export(module, example, path, input_name, output_name, training=False):
    from torch.onnx._internal.torchscript_exporter import onnx_proto_utils
    orig = onnx_proto_utils._add_onnxscript_fn
    onnx_proto_utils._add_onnxscript_fn = lambda proto, custom_opsets: proto  # needs the `onnx` package only for onnxscript functions
    buf = io.BytesIO()
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            torch.onnx.export(module, example, buf, dynamo=False, input_names=[input_name], output_names=[output_name], opset_version=17,
                              do_constant_folding=not training,
                              training=torch.onnx.TrainingMode.TRAINING if training else torch.onnx.TrainingMode.EVAL,
                              dynamic_axes={input_name: {0: "B"}} if input_name == "input" else {input_name: {0: "B", 1: "T"}})
    finally:
        onnx_proto_utils._add_onnxscript_fn = orig
    with open(path, "wb") as f:
        f.write(buf.getvalue())
    return path


class PyanNet(nn.Module):
    def __init__(self, w):
        super().__init__()
        T = lambda a: torch.from_numpy(np.ascontiguousarray(a, np.float32))
        self.wav_norm = nn.InstanceNorm1d(1, affine=True)
        self.conv0 = nn.Conv1d(1, 80, 251, stride=10, bias=False)
        self.conv1 = nn.Conv1d(80, 60, 5)
        self.conv2 = nn.Conv1d(60, 60, 5)
        self.norm0, self.norm1, self.norm2 = (nn.InstanceNorm1d(c, affine=True) for c in (80, 60, 60))
        self.lstm = nn.LSTM(60, 128, num_layers=4, bidirectional=True, batch_first=True)
        self.linear0, self.linear1, self.classifier = nn.Linear(256, 128), nn.Linear(128, 128), nn.Linear(128, 7)
        sd = {k: T(v) for k, v in w.items()}
        missing = self.load_state_dict(sd, strict=True)
        assert not missing.missing_keys and not missing.unexpected_keys

    def forward(self, x):
        x = self.wav_norm(x)
        x = TF.leaky_relu(self.norm0(TF.max_pool1d(self.conv0(x).abs(), 3, 3)))
        x = TF.leaky_relu(self.norm1(TF.max_pool1d(self.conv1(x), 3, 3)))
        x = TF.leaky_relu(self.norm2(TF.max_pool1d(self.conv2(x), 3, 3)))
        x, _ = self.lstm(x.transpose(1, 2))
        x = TF.leaky_relu(self.linear0(x))
        x = TF.leaky_relu(self.linear1(x))
        return torch.log_softmax(self.classifier(x), -1)


def export_pyannet_onnx(path, weights):
    """weights: oracle.pyannet.pyannet_weights(seed) (PyTorch names / layouts)."""
    m = PyanNet(weights).eval()
    return _export(m, torch.zeros(1, 1, 160000), path, "input", "output")


class _Block(nn.Module):
    def __init__(self, c_in, planes, stride):
        super().__init__()
        self.conv1 = nn.Conv2d(c_in, planes, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.shortcut = None
        if stride != 1 or c_in != planes:
            self.shortcut = nn.Sequential(nn.Conv2d(c_in, planes, 1, stride, bias=False), nn.BatchNorm2d(planes))

    def forward(self, x):
        y = torch.relu(self.bn1(self.conv1(x)))
        y = self.bn2(self.conv2(y))
        return torch.relu(y + (self.shortcut(x) if self.shortcut is not None else x))


class ResNet34(nn.Module):
    """wespeaker/models/resnet.py (published architecture): feats [B, T, 80] -> [B, 256]."""

    def __init__(self, seed):
        super().__init__()
        from oracle import weights as W
        self.conv1 = nn.Conv2d(1, 32, 3, 1, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(32)
        layers, c_in = [], 32
        for planes, blocks, stride in ((32, 3, 1), (64, 4, 2), (128, 6, 2), (256, 3, 2)):
            for bi in range(blocks):
                layers.append(_Block(c_in, planes, stride if bi == 0 else 1))
                c_in = planes
        self.blocks = nn.ModuleList(layers)
        self.seg_1 = nn.Linear(5120, 256)
        T = lambda a: torch.from_numpy(np.ascontiguousarray(a, np.float32))

        def load(conv, bn, name, ci, co, k):
            fan = ci * k * k
            tail = name.endswith("conv2") or name.endswith("shortcut")
            conv.weight.data = T(W.synth(seed, f"resnet34.{name}.weight", (co, ci, k, k), 0.0, np.sqrt(6.0 / fan), native_ok=False))
            bn.weight.data = T(W.synth(seed, f"resnet34.{name}.bn.weight", (co,), 0.7 if tail else 1.0, 0.1, native_ok=False))
            bn.bias.data = T(W.synth(seed, f"resnet34.{name}.bn.bias", (co,), 0.0, 0.1, native_ok=False))
            bn.running_mean.data = T(W.synth(seed, f"resnet34.{name}.bn.running_mean", (co,), 0.0, 0.1, native_ok=False))
            bn.running_var.data = T(W.synth(seed, f"resnet34.{name}.bn.running_var", (co,), 1.0, 0.2, native_ok=False))

        load(self.conv1, self.bn1, "conv1", 1, 32, 3)
        i, c_in = 0, 32
        for li, (planes, blocks, stride) in enumerate(((32, 3, 1), (64, 4, 2), (128, 6, 2), (256, 3, 2)), start=1):
            for bi in range(blocks):
                b = self.blocks[i]
                load(b.conv1, b.bn1, f"layer{li}.{bi}.conv1", c_in, planes, 3)
                load(b.conv2, b.bn2, f"layer{li}.{bi}.conv2", planes, planes, 3)
                if b.shortcut is not None:
                    load(b.shortcut[0], b.shortcut[1], f"layer{li}.{bi}.shortcut", c_in, planes, 1)
                c_in = planes
                i += 1
        self.seg_1.weight.data = T(W.synth(seed, "resnet34.seg_1.weight", (256, 5120), 0.0, 1.0 / np.sqrt(5120.0), native_ok=False))
        self.seg_1.bias.data = T(W.synth(seed, "resnet34.seg_1.bias", (256,), 0.0, 0.05, native_ok=False))

    def forward(self, feats):
        x = feats.permute(0, 2, 1).unsqueeze(1)  # [B, 1, 80, T]
        x = torch.relu(self.bn1(self.conv1(x)))
        for b in self.blocks:
            x = b(x)
        x = x.reshape(x.shape[0], x.shape[1] * x.shape[2], x.shape[3])  # index c * 10 + f
        stats = torch.cat([x.mean(-1), torch.sqrt(x.var(-1, unbiased=True) + 1e-7)], -1)
        return self.seg_1(stats)


def export_resnet34_onnx(path, seed=1234, fold_bn=True):
    m = ResNet34(seed)
    m.eval()
    if not fold_bn:  # a training-mode export traces a training-mode forward: momentum 0 keeps the running statistics as seeded
        for mod in m.modules():
            if isinstance(mod, nn.BatchNorm2d):
                mod.momentum = 0.0
    return _export(m, torch.zeros(1, 200, 80), path, "feats", "embs", training=not fold_bn)


def write_silero_ggml(path, w, use_f16=True):
    """w: oracle.vad weights dict ("enc.i.weight" [co, ci, 3], "enc.i.bias", "lstm.w_ih" ..., "final.weight" [128], "final.bias" [1])
    plus "basis" [258, 256].  Layout of whisper.cpp's convert-silero-vad-to-ggml.py: header, then tensor records (dims innermost
    first); conv weights in f16 when use_f16, everything else f32."""
    chans = ((129, 128), (128, 64), (64, 64), (64, 128))
    with open(path, "wb") as f:
        f.write(struct.pack("<I", 0x67676D6C))
        mt = b"silero-16k"
        f.write(struct.pack("<i", len(mt)))
        f.write(mt)
        f.write(struct.pack("<3i", 5, 1, 2))
        f.write(struct.pack("<i", 4))
        for ci, co in chans:
            f.write(struct.pack("<3i", ci, co, 3))
        f.write(struct.pack("<4i", 128, 128, 128, 1))

        def rec(name, a, f16=False):
            a = np.ascontiguousarray(a, np.float32)
            nb = name.encode()
            f.write(struct.pack("<3i", a.ndim, len(nb), 1 if f16 else 0))
            for d in reversed(a.shape):
                f.write(struct.pack("<i", d))
            f.write(nb)
            f.write((a.astype(np.float16) if f16 else a).tobytes())

        rec("_model.stft.forward_basis_buffer", np.asarray(w["basis"], np.float32).reshape(258, 1, 256))
        for i in range(4):
            rec(f"_model.encoder.{i}.reparam_conv.weight", w[f"enc.{i}.weight"], use_f16)
            rec(f"_model.encoder.{i}.reparam_conv.bias", w[f"enc.{i}.bias"])
        rec("_model.decoder.rnn.weight_ih", w["lstm.w_ih"])
        rec("_model.decoder.rnn.weight_hh", w["lstm.w_hh"])
        rec("_model.decoder.rnn.bias_ih", w["lstm.b_ih"])
        rec("_model.decoder.rnn.bias_hh", w["lstm.b_hh"])
        rec("_model.decoder.decoder.2.weight", np.asarray(w["final.weight"], np.float32).reshape(1, 128, 1))
        rec("_model.decoder.decoder.2.bias", w["final.bias"])
    return path
