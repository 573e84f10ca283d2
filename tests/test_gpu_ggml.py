"""GPU suite for SURVEY §8f row 2: a context loaded from a ggml-<model>.bin checkpoint (WhisperContext::new_with_params(model_path),
reference src/transcribe.rs:154) behaves exactly like the seeded context the file was written from — encoder output, greedy tokens,
token timestamps, DTW times and the text assembled from the FILE's vocabulary; the f16 container agrees within f16 rounding."""
import numpy as np
import pytest

from conftest import synth_audio
from ggml_writer import write_ggml

pytestmark = pytest.mark.gpu


def test_context_from_ggml_file_equals_seeded_context(wdr, oracle, tmp_path):
    from oracle import filters, vocab as V, weights as W
    arch_name = "tiny.en"
    arch = W.ARCHS[arch_name]
    w = W.whisper_weights(arch_name, 1234)
    toks = [V.token_text(i, arch["n_vocab"]).encode() for i in range(50256)]  # the checker's vocabulary: voice_length() of the token
    # strings feeds the heuristic timestamps, so parity with the checker needs the same strings
    path = str(tmp_path / "ggml-tiny.en.bin")
    write_ggml(path, arch, w, filters.whisper_mel_filters(80), toks, use_f16=False)
    pcm = synth_audio(31, 12.0)
    seeded = wdr.Context(arch_name, seed=1234, enable_dtw=True)
    loaded = wdr.Context(arch_name, enable_dtw=True, model_path=path)
    assert (loaded.dims.n_vocab, loaded.dims.n_audio_state, loaded.dims.n_text_layer) == (51864, 384, 4)
    s1, s2 = seeded.create_state(), loaded.create_state()
    win = np.zeros((1, 480000), np.int16)
    win[0, : len(pcm)] = pcm
    nv = np.array([len(pcm)], np.int32)
    h1, h2 = s1.encode_chunks(win, nv), s2.encode_chunks(win, nv)
    # the seeded context evaluates the sinusoidal positions / filterbank itself, the file carries numpy's: equal to fp32 noise
    assert np.abs(h1 - h2).max() <= 2e-3 * np.abs(h1).max(), "weights / filterbank / positions differ"
    x = pcm.astype(np.float32) / np.float32(32768.0)
    ref_h = oracle.whisper_encode(np.ascontiguousarray(oracle.log_mel(x, filters.whisper_mel_filters(80))[:, :3000]), arch_name, W.pack_encoder(arch_name, w))
    assert np.abs(h2[0] - ref_h).max() <= 1e-2 * np.abs(ref_h).max()
    # the decode of the loaded context against the oracle on the same weights (and the loaded context's own encoder output)
    from oracle import full
    dec = oracle.Decoder(arch_name, W.pack_decoder(arch_name, w), bf16=True)
    ref = full.full_window(dec, h2[0], x)
    dec.close()
    b = s2.full(pcm)
    assert len(b) == len(ref["segments"]) and len(b) >= 1
    for y, r in zip(b, ref["segments"]):
        assert [(t.id, t.t0, t.t1, t.t_dtw) for t in y["tokens"]] == [(t.id, t.t0, t.t1, t.t_dtw) for t in r["tokens"]]
        assert y["text"] == "".join(toks[t.id].decode() for t in y["tokens"] if t.id < 50256)
    # a checkpoint with other token strings: same ids, text assembled from the FILE's vocabulary (the " " token is found by its string)
    toks2 = [b"<" + t + b">" for t in toks]
    toks2[220], toks2[221] = toks2[221], b" "
    p2 = str(tmp_path / "ggml-tiny.en-vocab2.bin")
    write_ggml(p2, arch, w, filters.whisper_mel_filters(80), toks2, use_f16=False)
    lv = wdr.Context(arch_name, enable_dtw=True, model_path=p2)
    sv = lv.create_state()
    c = sv.full(pcm)
    assert c and all(seg["text"] == "".join(toks2[t.id].decode() for t in seg["tokens"] if t.id < 50256) for seg in c)
    assert "<" in c[0]["text"]
    sv.close()
    lv.close()
    # f16 container: same architecture, values within f16 rounding of the bf16 "checkpoint"
    p16 = str(tmp_path / "ggml-tiny.en-f16.bin")
    write_ggml(p16, arch, w, filters.whisper_mel_filters(80), toks, use_f16=True)
    l16 = wdr.Context(arch_name, enable_dtw=True, model_path=p16)
    s3 = l16.create_state()
    h3 = s3.encode_chunks(win, nv)
    assert np.abs(h2 - h3).max() <= 2e-2 * np.abs(h2).max()
    for s in (s1, s2, s3):
        s.close()
    for c in (seeded, loaded, l16):
        c.close()
    with pytest.raises(wdr.WdrError):
        wdr.Context(arch_name, model_path=str(tmp_path / "nope.bin"))
