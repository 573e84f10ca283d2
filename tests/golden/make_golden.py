#!/usr/bin/env python
"""Generates the golden vectors under tests/golden/ from INDEPENDENT implementations available in the build container
(`transformers` 5.5 OpenAI-lineage Whisper code, `torchaudio` Kaldi compliance, `scipy`), on seeded inputs.

The reference crate's own tests pin nothing on the hot path and its native dependencies (whisper.cpp, ONNX
Runtime, kaldi-native-fbank) are absent offline, so these vectors are the durable pins of the oracle: tests/test_golden.py checks
oracle/ against them on the CPU, tests/test_gpu_golden.py checks the CUDA path against them on the B200.

    python tests/golden/make_golden.py          # rewrites tests/golden/*.npz (small: < 1 MB in total)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from conftest import synth_audio  # noqa: E402


def mel():
    """WhisperFeatureExtractor log-mel (80 and 128 bins) of 1.5 s of seeded audio: the first 160 frames (frames past the audio
    are the constant floor).  Same framing as whisper.cpp for these frames (tests/test_oracle_signal.py)."""
    from transformers import WhisperFeatureExtractor
    pcm = synth_audio(101, 1.5)
    x = pcm.astype(np.float32) / np.float32(32768.0)
    out = {"pcm": pcm}
    for n_mel in (80, 128):
        fe = WhisperFeatureExtractor(feature_size=n_mel)
        out[f"mel{n_mel}"] = fe(x, sampling_rate=16000, return_tensors="np")["input_features"][0][:, :160].astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "mel.npz"), **out)


def median_dtw():
    """_median_filter (width 7, reflect) and _dynamic_time_warping (OpenAI tie-breaks) incl. ties, an inf cell, N > M."""
    from transformers.models.whisper.generation_whisper import _dynamic_time_warping, _median_filter
    rng = np.random.default_rng(202)
    w = rng.standard_normal((2, 9, 40)).astype(np.float32)
    out = {"med_in": w, "med_out": _median_filter(torch.from_numpy(w), 7).numpy()}
    cases = [rng.standard_normal((6, 25)).astype(np.float32), np.round(rng.standard_normal((5, 9)) * 2).astype(np.float32) / 2,
             rng.standard_normal((12, 5)).astype(np.float32), np.zeros((3, 4), np.float32)]
    cases[0][2, 7] = np.inf
    for i, x in enumerate(cases):
        ti, tj = _dynamic_time_warping(x.astype(np.float64))
        out[f"dtw{i}_x"], out[f"dtw{i}_ti"], out[f"dtw{i}_tj"] = x, ti.astype(np.int32), tj.astype(np.int32)
    out["n_dtw"] = np.int32(len(cases))
    np.savez_compressed(os.path.join(HERE, "median_dtw.npz"), **out)


def fbank():
    """torchaudio.compliance.kaldi.fbank with kaldi-native-fbank's options on 0.6 s of int16-scale audio."""
    import torchaudio.compliance.kaldi as k
    pcm = synth_audio(303, 0.6)
    ref = k.fbank(torch.from_numpy(pcm.astype(np.float32))[None], num_mel_bins=80, frame_length=25, frame_shift=10, dither=0.0, energy_floor=0.0,
                  sample_frequency=16000, window_type="povey", preemphasis_coefficient=0.97, remove_dc_offset=True, snip_edges=True,
                  low_freq=20, high_freq=0, use_log_fbank=True, use_power=True).numpy()
    np.savez_compressed(os.path.join(HERE, "fbank.npz"), pcm=pcm, fbank=ref.astype(np.float32))


def encoder():
    """transformers WhisperEncoder (tiny.en geometry, the seeded synthetic weights of oracle/weights.py) on a seeded log-mel window:
    a [1500, 384] hidden state is 2.3 MB, so the fixture keeps rows 0, 1, 749, 1499 and per-row L2 norms of all rows."""
    from transformers import WhisperConfig
    from transformers.models.whisper.modeling_whisper import WhisperEncoder
    from oracle import filters, native, weights as W
    arch = "tiny.en"
    a = W.ARCHS[arch]
    w = W.whisper_weights(arch, seed=1234)
    cfg = WhisperConfig(d_model=a["d"], encoder_layers=a["n_enc"], encoder_attention_heads=a["n_head"], encoder_ffn_dim=4 * a["d"],
                        num_mel_bins=a["n_mel"], activation_function="gelu_pytorch_tanh", max_source_positions=1500)
    enc = WhisperEncoder(cfg).eval()
    sd = {}
    m = {"conv1.weight": "encoder.conv1.weight", "conv1.bias": "encoder.conv1.bias", "conv2.weight": "encoder.conv2.weight",
         "conv2.bias": "encoder.conv2.bias", "embed_positions.weight": "encoder.positional_embedding",
         "layer_norm.weight": "encoder.ln_post.weight", "layer_norm.bias": "encoder.ln_post.bias"}
    for k2, v in m.items():
        sd[k2] = torch.from_numpy(w[v])
    for l in range(a["n_enc"]):
        p, q = f"layers.{l}.", f"encoder.blocks.{l}."
        for hf, oa in (("self_attn.q_proj", "attn.query"), ("self_attn.k_proj", "attn.key"), ("self_attn.v_proj", "attn.value"),
                       ("self_attn.out_proj", "attn.out"), ("fc1", "mlp.0"), ("fc2", "mlp.2"), ("self_attn_layer_norm", "attn_ln"),
                       ("final_layer_norm", "mlp_ln")):
            sd[p + hf + ".weight"] = torch.from_numpy(w[q + oa + ".weight"])
            if q + oa + ".bias" in w:
                sd[p + hf + ".bias"] = torch.from_numpy(w[q + oa + ".bias"])
    missing, unexpected = enc.load_state_dict(sd, strict=False)
    assert not unexpected and all("k_proj.bias" in x for x in missing), (missing, unexpected)
    pcm = synth_audio(404, 4.0)
    x = pcm.astype(np.float32) / np.float32(32768.0)
    mel = np.ascontiguousarray(native.log_mel(x, filters.whisper_mel_filters(80))[:, :3000])
    with torch.no_grad():
        h = enc(torch.from_numpy(mel)[None]).last_hidden_state[0].numpy()
    rows = np.array([0, 1, 749, 1499])
    np.savez_compressed(os.path.join(HERE, "encoder_tiny_en.npz"), pcm=pcm, rows=rows, hidden_rows=h[rows].astype(np.float32),
                        row_norms=np.linalg.norm(h, axis=1).astype(np.float32))


def clustering():
    """scipy average-linkage clustering of a seeded cosine-similarity matrix cut at distance 1 - threshold (labels in
    first-appearance order), and a hand-scanned leader clustering (strict >, cap -> best match)."""
    from scipy.cluster.hierarchy import fcluster, linkage
    from scipy.spatial.distance import squareform
    rng = np.random.default_rng(505)
    centers = rng.standard_normal((4, 32))
    emb = np.concatenate([c + 0.35 * rng.standard_normal((6, 32)) for c in centers]).astype(np.float32)
    emb = emb[rng.permutation(len(emb))]
    n = emb / np.linalg.norm(emb, axis=1, keepdims=True)
    S = (n @ n.T).astype(np.float32)
    D = 1.0 - S.astype(np.float64)
    np.fill_diagonal(D, 0.0)
    Z = linkage(squareform((D + D.T) / 2, checks=False), method="average")
    raw = fcluster(Z, t=1.0 - 0.5 - 1e-9, criterion="distance")  # merge while similarity > 0.5 (strict)
    remap, labels = {}, []
    for r in raw:
        remap.setdefault(int(r), len(remap) + 1)  # ids start at 1, in order of first member
        labels.append(remap[int(r)])
    np.savez_compressed(os.path.join(HERE, "clustering.npz"), emb=emb, S=S, agglomerative_thr05=np.array(labels, np.int32))


def formatter():
    """The reference's OWN fixture: /root/reference/segments.json is the output of `process_segments` that examples/test.rs wrote
    (FormattingOverrides max_chars_per_line 20, max_lines 2; examples/test.rs:36-50).  Copied verbatim as data (cues with their
    word timestamps) — the only artefact in the reference that pins anything near the hot path."""
    import json
    src = "/root/reference/segments.json"
    cues = json.load(open(src))
    json.dump(dict(source="tmoroney/whisper-diarize-rs segments.json (written by examples/test.rs)", overrides=dict(max_chars_per_line=20, max_lines=2),
                   language="en", cues=cues), open(os.path.join(HERE, "formatter_segments.json"), "w"), indent=0)


if __name__ == "__main__":
    for fn in (mel, median_dtw, fbank, encoder, clustering, formatter):
        fn()
        print("wrote", fn.__name__)
