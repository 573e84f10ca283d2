"""CPU suite for SURVEY §8f row 2, the other half: the dependency-free readers of the ONNX files (segmentation-3.0, WeSpeaker ResNet34;
reference src/engine.rs:90-91) and of ggml-silero-v5.1.2.bin (src/model_manager.rs:305-315) parse what torch's own ONNX exporter /
the converter layout write, extract every parameter in the layout the kernels upload (LSTM gates re-ordered from ONNX i,o,f,c to
i,f,g,o; Linear weights back to [out][in]; BatchNorm folded), and refuse files of another architecture or damaged files — without a GPU."""
import numpy as np
import pytest

import model_writers as MW


@pytest.fixture(scope="module")
def seg_file(tmp_path_factory):
    from oracle import pyannet as P
    w = P.pyannet_weights(1234)
    return MW.export_pyannet_onnx(str(tmp_path_factory.mktemp("onnx") / "segmentation-3.0.onnx"), w), w


def test_pyannet_onnx_parameters_are_extracted_exactly(wdr, seg_file):
    path, w = seg_file
    info = wdr.onnx_probe(path, wdr.ONNX_PYANNET)
    assert info["inputs"] == 1 and info["outputs"] == 1 and info["params"] == len(w) and info["nodes"] > 20
    for name, ref in w.items():  # raw_data initializers: bit-exact, in PyTorch layout (gate order restored)
        got = wdr.onnx_read_param(path, wdr.ONNX_PYANNET, name)
        assert got.shape == (ref.size,) and np.array_equal(got, np.asarray(ref, np.float32).ravel()), name
    with pytest.raises(wdr.WdrError):
        wdr.onnx_probe(path, wdr.ONNX_RESNET34)  # a PyanNet is not a ResNet34


def test_resnet34_onnx_folded_and_unfolded_batchnorm(wdr, tmp_path):
    from oracle import resnet as R
    ref = R.resnet_weights(1234)  # folded, conv kernels rounded to bf16
    for fold in (True, False):
        path = MW.export_resnet34_onnx(str(tmp_path / f"wespeaker_resnet34_{int(fold)}.onnx"), 1234, fold_bn=fold)
        info = wdr.onnx_probe(path, wdr.ONNX_RESNET34)
        assert info["emb_dim"] == 256 and info["params"] == 2 * 36 + 2
        for name, _, co, _, _ in R.conv_specs():
            wf, bf = ref[name]
            got_w = wdr.onnx_read_param(path, wdr.ONNX_RESNET34, name + ".weight").reshape(wf.shape)
            got_b = wdr.onnx_read_param(path, wdr.ONNX_RESNET34, name + ".bias")
            # the exporter / the reader fold in fp32; the oracle additionally rounds the kernel to bf16 (2^-8 relative)
            assert np.abs(got_w - wf).max() <= 2 ** -8 * np.abs(wf).max() + 1e-6, (fold, name)
            assert np.abs(got_b - bf).max() <= 1e-5 * max(1.0, np.abs(bf).max()), (fold, name)
        lw, lb = ref["seg_1"]
        assert np.array_equal(wdr.onnx_read_param(path, wdr.ONNX_RESNET34, "seg_1.weight").reshape(lw.shape), lw)
        assert np.array_equal(wdr.onnx_read_param(path, wdr.ONNX_RESNET34, "seg_1.bias"), lb)
        with pytest.raises(wdr.WdrError) as e:
            wdr.onnx_probe(path, wdr.ONNX_PYANNET)
        assert "PyanNet" in str(e.value)


def test_damaged_onnx_files_are_refused(wdr, seg_file, tmp_path):
    path, _ = seg_file
    raw = open(path, "rb").read()
    for name, data in (("trunc.onnx", raw[: len(raw) // 2]), ("garbage.onnx", bytes(range(256)) * 64), ("empty.onnx", b"")):
        p = str(tmp_path / name)
        open(p, "wb").write(data)
        with pytest.raises(wdr.WdrError):
            wdr.onnx_probe(p, wdr.ONNX_PYANNET)
    with pytest.raises(wdr.WdrError):
        wdr.onnx_probe(str(tmp_path / "missing.onnx"), wdr.ONNX_PYANNET)


def test_silero_ggml_header_and_index(wdr, tmp_path):
    from oracle import vad as V
    w = dict(V.vad_weights(1234))
    w["basis"] = w["stft.basis"]
    path = MW.write_silero_ggml(str(tmp_path / "ggml-silero-v5.1.2.bin"), w)
    info = wdr.silero_probe(path)
    assert info["version"] == (5, 1, 2) and info["n_encoder_layers"] == 4 and info["n_tensors"] == 15
    assert info["encoder"] == [(129, 128, 3), (128, 64, 3), (64, 64, 3), (64, 128, 3)]
    assert (info["lstm_input"], info["lstm_hidden"], info["final_in"], info["final_out"]) == (128, 128, 128, 1)
    raw = open(path, "rb").read()
    bad = str(tmp_path / "trunc.bin")
    open(bad, "wb").write(raw[: len(raw) - 100])
    with pytest.raises(wdr.WdrError):
        wdr.silero_probe(bad)
    bad2 = str(tmp_path / "magic.bin")
    open(bad2, "wb").write(b"GGUF" + raw[4:])
    with pytest.raises(wdr.WdrError):
        wdr.silero_probe(bad2)
