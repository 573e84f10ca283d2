"""GPU tests for SURVEY §8f row 2: the models LOADED FROM FILES equal the seeded models and the oracle — segmentation-3.0.onnx
(pyannote_rs::get_segments, reference src/engine.rs:90, 117-122), the WeSpeaker ResNet34 ONNX export (EmbeddingExtractor::new,
src/engine.rs:91, src/transcribe.rs:343) and ggml-silero-v5.1.2.bin (WhisperVadContext::new, src/vad.rs:15-17).  The files are
written by tests/model_writers.py (torch's own ONNX exporter / the converter's ggml layout) from the oracle's seeded tensors, so a
context loaded from a file must reproduce the context created without one (bit for bit where the file holds the same fp32 values)."""
import numpy as np
import pytest

import model_writers as MW
from conftest import synth_audio

pytestmark = pytest.mark.gpu


def test_segmenter_from_onnx_equals_seeded_and_oracle(wdr, tmp_path):
    from oracle import pyannet as P
    w = P.pyannet_weights(1234)
    path = MW.export_pyannet_onnx(str(tmp_path / "segmentation-3.0.onnx"), w)
    pcm = synth_audio(61, 17.3, n_speakers=2)
    a = wdr.Segmenter(seed=1234)
    b = wdr.Segmenter(path=path)
    sa, sb = a.scores(pcm), b.scores(pcm)
    assert np.array_equal(sa, sb)  # identical fp32 parameters -> identical kernels -> identical scores
    ga, gb = a.get_segments(pcm), b.get_segments(pcm)
    assert [(g["start"], g["end"], g["i0"], g["i1"]) for g in ga] == [(g["start"], g["end"], g["i0"], g["i1"]) for g in gb] and len(gb) >= 1
    ref = P.pyannet_forward(np.pad(pcm, (0, 2 * P.WINDOW - len(pcm)))[: P.WINDOW].astype(np.float32), w)
    assert np.abs(sb[0] - ref).max() < 5e-3
    a.close()
    b.close()
    with pytest.raises(wdr.WdrError):
        wdr.Segmenter(path=str(tmp_path / "missing.onnx"))


def test_embedding_from_onnx_equals_oracle(wdr, oracle, tmp_path):
    from oracle import resnet
    rw = resnet.resnet_weights(1234)
    pcm = synth_audio(62, 3.1)
    ref = resnet.compute(pcm, rw, oracle.kaldi_fbank)
    seeded = wdr.EmbeddingExtractor(seed=1234)
    e0 = seeded.compute(pcm)
    seeded.close()
    for fold in (True, False):
        path = MW.export_resnet34_onnx(str(tmp_path / f"wespeaker_resnet34_{int(fold)}.onnx"), 1234, fold_bn=fold)
        ex = wdr.EmbeddingExtractor(path=path)
        assert ex.dim == 256
        e = ex.compute(pcm)
        ex.close()
        cos = float(e @ ref / (np.linalg.norm(e) * np.linalg.norm(ref)))
        assert cos >= 0.9995 and np.abs(e - ref).max() <= 3e-2 * np.abs(ref).max(), (fold, cos)
        # against the seeded context: the only difference is where the fp32 fold was rounded to bf16
        cos0 = float(e @ e0 / (np.linalg.norm(e) * np.linalg.norm(e0)))
        assert cos0 >= 0.9999, (fold, cos0)
    # a PyanNet file is not an embedding model
    from oracle import pyannet as P
    other = MW.export_pyannet_onnx(str(tmp_path / "seg.onnx"), P.pyannet_weights(1234))
    with pytest.raises(wdr.WdrError) as err:
        wdr.EmbeddingExtractor(path=other)
    assert "ResNet34" in str(err.value)


def test_vad_from_silero_ggml(wdr, tmp_path):
    from oracle import vad as V, weights as W
    w = dict(V.vad_weights(1234))
    w["basis"] = w["stft.basis"]
    pcm = synth_audio(63, 8.0, n_speakers=2)
    x = pcm.astype(np.float32) / np.float32(32768.0)
    seeded = wdr.VadContext(seed=1234)
    p0 = seeded.detect_speech(x)
    seeded.close()
    # f32 file: the same parameters as the seeded context (its STFT basis is the true windowed DFT basis the oracle stores)
    v32 = wdr.VadContext(path=MW.write_silero_ggml(str(tmp_path / "silero_f32.bin"), w, use_f16=False))
    p32 = v32.detect_speech(x)
    assert np.abs(p32 - p0).max() < 1e-6
    assert np.abs(p32 - V.silero_probs(x, V.vad_weights(1234))).max() < 2e-4
    segs = v32.segments_from_samples(x, wdr.vad_default_params(min_silence_duration_ms=100))
    assert segs == V.segments_from_probs(p32, dict(min_silence_duration_ms=100))
    v32.close()
    # the converter's layout: conv kernels in f16 -> the oracle run on the same f16-rounded kernels
    v16 = wdr.VadContext(path=MW.write_silero_ggml(str(tmp_path / "ggml-silero-v5.1.2.bin"), w, use_f16=True))
    w16 = dict(V.vad_weights(1234))
    for i in range(4):
        w16[f"enc.{i}.weight"] = w16[f"enc.{i}.weight"].astype(np.float16).astype(np.float32)
    p16 = v16.detect_speech(x)
    assert np.abs(p16 - V.silero_probs(x, w16)).max() < 2e-4
    v16.close()
    with pytest.raises(wdr.WdrError):
        wdr.VadContext(path=str(tmp_path / "silero_missing.bin"))
