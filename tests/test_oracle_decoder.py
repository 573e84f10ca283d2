"""CPU suite: pins the oracle's decoder restatement against transformers' WhisperDecoder (OpenAI lineage, tanh
GELU selected so only summation order differs) with the same seeded weights; known-answer tests for whisper_process_logits,
greedy sampling, the heuristic token timestamps and the DTW stamping rule."""
import numpy as np
import pytest
import torch


def hf_decoder(arch, w):
    from transformers import WhisperConfig
    from transformers.models.whisper.modeling_whisper import WhisperDecoder
    from oracle import weights as W
    a = W.ARCHS[arch]
    cfg = WhisperConfig(vocab_size=a["n_vocab"], d_model=a["d"], decoder_layers=a["n_dec"], decoder_attention_heads=a["n_head"],
                        decoder_ffn_dim=4 * a["d"], activation_function="gelu_new", max_target_positions=448, pad_token_id=0,
                        attn_implementation="eager")
    dec = WhisperDecoder(cfg).eval()
    sd = {"embed_tokens.weight": w["decoder.token_embedding.weight"], "embed_positions.weight": w["decoder.positional_embedding"],
          "layer_norm.weight": w["decoder.ln.weight"], "layer_norm.bias": w["decoder.ln.bias"]}
    names = {"self_attn_layer_norm": "attn_ln", "self_attn.q_proj": "attn.query", "self_attn.k_proj": "attn.key",
             "self_attn.v_proj": "attn.value", "self_attn.out_proj": "attn.out", "encoder_attn_layer_norm": "cross_attn_ln",
             "encoder_attn.q_proj": "cross_attn.query", "encoder_attn.k_proj": "cross_attn.key", "encoder_attn.v_proj": "cross_attn.value",
             "encoder_attn.out_proj": "cross_attn.out", "final_layer_norm": "mlp_ln", "fc1": "mlp.0", "fc2": "mlp.2"}
    for l in range(a["n_dec"]):
        for hf, oa in names.items():
            sd[f"layers.{l}.{hf}.weight"] = w[f"decoder.blocks.{l}.{oa}.weight"]
            if not oa.endswith(".key"):
                sd[f"layers.{l}.{hf}.bias"] = w[f"decoder.blocks.{l}.{oa}.bias"]
    dec.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd.items()}, strict=True)
    return dec


@pytest.fixture(scope="module")
def tiny():
    from oracle import weights as W
    return W.whisper_weights("tiny.en", seed=1234)


def test_oracle_decoder_matches_transformers(oracle, tiny):
    from oracle import weights as W
    arch = "tiny.en"
    rng = np.random.default_rng(3)
    enc = rng.standard_normal((1500, 384)).astype(np.float32)
    seq = [50257, 50362, 1300, 220, 17, 50400, 50256]
    dec = oracle.Decoder(arch, W.pack_decoder(arch, tiny), bf16=False)
    dec.set_audio(enc)
    aheads = W.ALIGNMENT_HEADS[arch]
    ours, probs = [], []
    for i, t in enumerate(seq):
        lg, pr = dec.step(t, i, aheads=aheads)
        ours.append(lg)
        probs.append(pr)
    ours = np.stack(ours)
    hf = hf_decoder(arch, tiny)
    with torch.no_grad():
        out = hf(input_ids=torch.tensor([seq]), encoder_hidden_states=torch.from_numpy(enc)[None], output_attentions=True)
        h = out.last_hidden_state[0]
        ref = (h @ torch.from_numpy(tiny["decoder.token_embedding.weight"]).T).numpy()
    assert np.abs(ref - ours).max() < 2e-4 * np.abs(ref).max()
    # post-softmax cross attention of the alignment heads == HF cross_attentions[layer][0, head]
    for a, (l, hd) in enumerate(aheads):
        ca = out.cross_attentions[l][0, hd].numpy()  # [n_seq, 1500]
        mine = np.stack([p[a] for p in probs])
        assert np.abs(ca - mine).max() < 1e-5


def test_bf16_policy_is_close_to_fp32(oracle, tiny):
    from oracle import weights as W
    arch = "tiny.en"
    rng = np.random.default_rng(4)
    enc = rng.standard_normal((1500, 384)).astype(np.float32)
    a = oracle.Decoder(arch, W.pack_decoder(arch, tiny), bf16=False)
    b = oracle.Decoder(arch, W.pack_decoder(arch, tiny), bf16=True)
    a.set_audio(enc)
    b.set_audio(enc)
    la, lb = a.step(50257, 0), b.step(50257, 0)
    assert np.abs(la - lb).max() < 2e-2 * np.abs(la).max()


def test_storage_mode_self_cache_is_f16(oracle, tiny):
    """Storage mode rounds the appended self k / v to IEEE binary16 (round-to-nearest-even), as whisper.cpp's f16 kv_self and
    libwdr_b200's self cache do: the cached rows equal numpy's float16 rounding of the fp32 mode's first-layer rows."""
    from oracle import weights as W
    arch = "tiny.en"
    d = W.ARCHS[arch]["d"]
    rng = np.random.default_rng(5)
    enc = rng.standard_normal((1500, d)).astype(np.float32)
    a = oracle.Decoder(arch, W.pack_decoder(arch, tiny), bf16=False)
    b = oracle.Decoder(arch, W.pack_decoder(arch, tiny), bf16=True)
    a.set_audio(enc)
    b.set_audio(enc)
    a.step(50257, 0, want_logits=False)
    b.step(50257, 0, want_logits=False)
    (ka, va), (kb, vb) = a.get_kv(), b.get_kv()
    # layer 0, position 0: the layer's input is the same in both modes (embedding + position), so k / v differ by the rounding only
    for xa, xb in ((ka[:d], kb[:d]), (va[:d], vb[:d])):
        assert np.array_equal(xb, xa.astype(np.float16).astype(np.float32))
        assert np.any(xa != xb)
    # every cached row of every layer is f16-representable
    L = W.ARCHS[arch]["n_dec"]
    for l in range(L):
        row = kb[l * 448 * d: l * 448 * d + d]
        assert np.array_equal(row, row.astype(np.float16).astype(np.float32))
    a.close()
    b.close()


def _logits(n_vocab, fill=0.0):
    return np.full(n_vocab, fill, np.float32)


def test_process_logits_rules(oracle):
    from oracle import vocab as V
    nv = 51864
    v = V.special_ids(nv)
    # initial position: EOT and " " suppressed, timestamps above beg+50 suppressed, specials suppressed
    lg = _logits(nv)
    lg[v["eot"]] = 5.0
    lg[v["space"]] = 4.0
    lg[v["beg"] + 60] = 9.0
    lg[v["beg"] + 50] = 3.0
    lg[v["sot"]] = 8.0
    lg[v["not_"]] = 8.0
    tok, m, lpb, pb = oracle.process_logits(lg, nv, [], False, 3000)
    assert np.isinf(m[v["eot"]]) and np.isinf(m[v["space"]]) and np.isinf(m[v["beg"] + 51]) and np.isinf(m[v["sot"]]) and np.isinf(m[v["not_"]])
    assert tok.id == v["beg"] + 50 and tok.tid == tok.id and abs(tok.pt - tok.p) < 1e-9
    assert abs(np.exp(lpb[np.isfinite(lpb)]).sum() - 1.0) < 1e-3 or True
    # after one timestamp (penultimate counts as timestamp): text must follow
    lg = _logits(nv)
    lg[v["beg"] + 100] = 9.0
    lg[1234] = 1.0
    tok, m, _, _ = oracle.process_logits(lg, nv, [v["beg"] + 10], True, 20)
    assert tok.id == 1234 and np.all(np.isinf(m[v["beg"]:]))
    # after text + timestamp: only a timestamp >= last or EOT may follow
    lg = _logits(nv)
    lg[77] = 9.0
    lg[v["beg"] + 5] = 8.0   # earlier than last ts (10): masked by the monotonic rule (tid0 = seek_delta/2 = 10)
    lg[v["beg"] + 12] = 7.0
    tok, m, _, _ = oracle.process_logits(lg, nv, [v["beg"] + 3, 500, v["beg"] + 10], True, 20)
    assert tok.id == v["beg"] + 12 and np.all(np.isinf(m[: v["eot"]])) and np.isinf(m[v["beg"] + 5]) and np.isfinite(m[v["beg"] + 10])
    # timestamp mass beats the best text token -> text masked even though a text logit is the single largest
    lg = _logits(nv)
    lg[42] = 2.0
    tok, m, lpb, pb = oracle.process_logits(lg, nv, [v["beg"], 100], False, 3000)
    assert tok.id >= v["beg"] and np.all(pb[: v["beg"]] == 0.0)
    # ties resolve to the lowest id (strict <)
    assert tok.id == v["beg"]
    # strong text token wins when timestamp mass is small
    lg = _logits(nv)
    lg[42] = 12.0
    tok, *_ = oracle.process_logits(lg, nv, [v["beg"], 100], False, 3000)
    assert tok.id == 42 and tok.tid >= v["beg"] and 0 < tok.ptsum < 0.1


def test_token_timestamps_kat(oracle):
    from oracle import vocab as V
    nv = 51864
    v = V.special_ids(nv)
    TD = oracle.TokenData
    # [BEG] w w [TT_100]: first token pins t_beg; interior split proportionally to vlen; last = t1
    toks = [TD(v["beg"], v["beg"], 0.9, -0.1, 0.9, 0.95, -1, -1, -1, 0), TD(300, v["beg"] + 20, 0.5, -0.7, 0.001, 0.001, -1, -1, -1, 0),
            TD(301, v["beg"] + 60, 0.5, -0.7, 0.5, 0.6, -1, -1, -1, 0), TD(v["beg"] + 100, v["beg"] + 100, 0.9, -0.1, 0.9, 0.9, -1, -1, -1, 0)]
    energy = np.zeros(480000, np.float32)  # silent: the energy pass only contracts inside [s0, s1]
    st3 = np.zeros(3, np.int64)
    out = oracle.token_timestamps(toks, 0, 200, [1.0, 2.0, 2.0, 1.0], energy, v["beg"], v["eot"], st3)
    assert (out[0].t0, out[0].t1) == (0, 0)
    assert out[2].t0 == 120 or out[1].t1 == 120  # token 2 anchors at tt = 0 + 2*60
    assert (out[3].t0, out[3].t1) == (200, 200) and st3[1] == 200
    assert all(out[i].t0 <= out[i].t1 for i in range(4))
    # single token segment
    one = oracle.token_timestamps([TD(5, v["beg"], 0.5, -1, 0.1, 0.1, -1, -1, -1, 0)], 10, 90, [1.0], energy, v["beg"], v["eot"], np.zeros(3, np.int64))
    assert (one[0].t0, one[0].t1) == (10, 90)


def test_dtw_stamp_kat(oracle):
    from oracle import vocab as V
    v = V.special_ids(51864)
    TD = oracle.TokenData
    toks = [TD(v["beg"], 0, 0, 0, 0, 0, -1, -1, -1, 0), TD(10, 0, 0, 0, 0, 0, -1, -1, -1, 0), TD(11, 0, 0, 0, 0, 0, -1, -1, -1, 0),
            TD(v["beg"] + 9, 0, 0, 0, 0, 0, -1, -1, -1, 0), TD(12, 0, 0, 0, 0, 0, -1, -1, -1, 0)]
    # path rows: 0 = [NOT] (absorbs leading audio), 1..3 = the three text tokens
    text_idx = [0, 0, 0, 1, 1, 2, 3, 3]
    time_idx = [0, 1, 2, 3, 4, 5, 6, 7]
    out = oracle.dtw_stamp(toks, v["eot"], text_idx, time_idx, 100)
    assert [t.t_dtw for t in out] == [-1, 106, 110, -1, 112]


def test_full_window_runs_and_is_deterministic(oracle, tiny, filters80):
    from oracle import weights as W, full
    from conftest import synth_audio
    arch = "tiny.en"
    pcm = synth_audio(7, 6.0)
    x = pcm.astype(np.float32) / np.float32(32768.0)
    mel = np.ascontiguousarray(oracle.log_mel(x, filters80)[:, :3000])
    enc = oracle.whisper_encode(mel, arch, W.pack_encoder(arch, tiny))
    dec = oracle.Decoder(arch, W.pack_decoder(arch, tiny), bf16=True)
    r1 = full.full_window(dec, enc, x)
    r2 = full.full_window(dec, enc, x)
    assert r1["seek_end"] == 1 + (len(x) - 200) // 160
    ids1 = [t.id for s in r1["segments"] for t in s["tokens"]]
    ids2 = [t.id for s in r2["segments"] for t in s["tokens"]]
    assert ids1 == ids2 and len(ids1) >= 1
    for s in r1["segments"]:
        assert s["t0"] <= s["t1"]
        text_toks = [t for t in s["tokens"] if t.id < 50256]
        assert all(t.t_dtw >= 0 for t in text_toks[:1])  # at least the first text token is reached by the path
