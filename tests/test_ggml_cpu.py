"""CPU suite for SURVEY §8f row 2: the ggml-<model>.bin reader (csrc/ggml_file.cu) parses what tests/ggml_writer.py writes — header,
vocabulary and tensor index — and refuses damaged / quantised files, without touching a GPU."""
import os
import struct

import numpy as np
import pytest

from ggml_writer import write_ggml


def _tiny_file(tmp_path, use_f16=False, name="ggml-tiny.en.bin"):
    from oracle import filters, vocab as V, weights as W
    arch = W.ARCHS["tiny.en"]
    w = W.whisper_weights("tiny.en", 1234)
    toks = [V.token_text(i, arch["n_vocab"]).encode() for i in range(50256)]
    p = str(tmp_path / name)
    write_ggml(p, arch, w, filters.whisper_mel_filters(80), toks, use_f16)
    return p, w


def test_probe_reads_header_vocab_and_index(wdr, tmp_path):
    p, w = _tiny_file(tmp_path)
    info = wdr.ggml_probe(p)
    assert (info["n_vocab"], info["n_audio_state"], info["n_audio_head"], info["n_audio_layer"], info["n_text_layer"], info["n_mels"]) == (51864, 384, 6, 4, 4, 80)
    assert info["n_audio_ctx"] == 1500 and info["n_text_ctx"] == 448 and info["ftype"] == 0
    assert info["n_tensors"] == len(w) and info["n_tokens"] == 50256
    p16, _ = _tiny_file(tmp_path, True, "ggml-tiny.en-f16.bin")
    assert wdr.ggml_probe(p16)["ftype"] == 1 and os.path.getsize(p16) < 0.6 * os.path.getsize(p)


def test_damaged_and_quantised_files_are_refused(wdr, tmp_path):
    p, _ = _tiny_file(tmp_path)
    raw = open(p, "rb").read()
    bad = str(tmp_path / "bad_magic.bin")
    open(bad, "wb").write(b"XXXX" + raw[4:])
    with pytest.raises(wdr.WdrError):
        wdr.ggml_probe(bad)
    trunc = str(tmp_path / "trunc.bin")
    open(trunc, "wb").write(raw[: len(raw) - 1000])
    with pytest.raises(wdr.WdrError):
        wdr.ggml_probe(trunc)
    # flip the first tensor's type field to a quantised type (q4_0 = 2)
    from oracle import vocab as V
    off = 4 + 44 + 8 + 80 * 201 * 4 + 4 + sum(4 + len(V.token_text(i, 51864).encode()) for i in range(50256))
    assert struct.unpack_from("<i", raw, off)[0] == 3  # n_dims of encoder.conv1.weight
    q = bytearray(raw)
    struct.pack_into("<i", q, off + 8, 2)
    qp = str(tmp_path / "quant.bin")
    open(qp, "wb").write(bytes(q))
    with pytest.raises(wdr.WdrError) as e:
        wdr.ggml_probe(qp)
    assert "quantised" in str(e.value)
    with pytest.raises(wdr.WdrError):
        wdr.ggml_probe(str(tmp_path / "missing.bin"))
