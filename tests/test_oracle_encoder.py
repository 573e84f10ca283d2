"""CPU suite: pins the oracle's encoder restatement against transformers' WhisperEncoder (OpenAI
lineage, tanh GELU selected so only summation order differs) with the same seeded weights, and the seeded weight
generator's determinism."""
import numpy as np
import torch


def hf_encoder(arch, w):
    from transformers import WhisperConfig
    from transformers.models.whisper.modeling_whisper import WhisperEncoder
    from oracle import weights as W
    a = W.ARCHS[arch]
    cfg = WhisperConfig(d_model=a["d"], encoder_layers=a["n_enc"], encoder_attention_heads=a["n_head"], encoder_ffn_dim=4 * a["d"],
                        num_mel_bins=a["n_mel"], activation_function="gelu_new", max_source_positions=1500)
    enc = WhisperEncoder(cfg).eval()
    sd = {"conv1.weight": w["encoder.conv1.weight"], "conv1.bias": w["encoder.conv1.bias"], "conv2.weight": w["encoder.conv2.weight"],
          "conv2.bias": w["encoder.conv2.bias"], "embed_positions.weight": w["encoder.positional_embedding"],
          "layer_norm.weight": w["encoder.ln_post.weight"], "layer_norm.bias": w["encoder.ln_post.bias"]}
    names = {"self_attn_layer_norm": "attn_ln", "self_attn.q_proj": "attn.query", "self_attn.k_proj": "attn.key",
             "self_attn.v_proj": "attn.value", "self_attn.out_proj": "attn.out", "final_layer_norm": "mlp_ln", "fc1": "mlp.0", "fc2": "mlp.2"}
    for l in range(a["n_enc"]):
        for hf, oa in names.items():
            sd[f"layers.{l}.{hf}.weight"] = w[f"encoder.blocks.{l}.{oa}.weight"]
            if oa != "attn.key":
                sd[f"layers.{l}.{hf}.bias"] = w[f"encoder.blocks.{l}.{oa}.bias"]
    enc.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd.items()}, strict=True)
    return enc


def test_oracle_encoder_matches_transformers(oracle, filters80):
    from oracle import weights as W
    arch = "tiny.en"
    w = W.whisper_weights(arch, seed=1234)
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(480000) * 0.05).astype(np.float32)
    mel = oracle.log_mel(x, filters80)[:, :3000]
    h = oracle.whisper_encode(mel, arch, W.pack_encoder(arch, w))
    with torch.no_grad():
        ref = hf_encoder(arch, w)(torch.from_numpy(mel)[None]).last_hidden_state[0].numpy()
    assert np.abs(ref - h).max() < 1e-4 * np.abs(ref).max()


def test_weight_generator_is_deterministic_and_bf16():
    from oracle import weights as W
    a = W.synth(1234, "encoder.blocks.0.attn.query.weight", (8, 8), bf16=True)
    b = W.synth(1234, "encoder.blocks.0.attn.query.weight", (8, 8), bf16=True)
    c = W.synth(1235, "encoder.blocks.0.attn.query.weight", (8, 8), bf16=True)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert np.all((a.view(np.uint32) & 0xFFFF) == 0)  # bf16-representable
    assert np.abs(a).max() <= W.W_SCALE * 1.01
    big = W.synth(7, "x", (200000,))
    assert abs(big.std() - 0.02) < 5e-4 and abs(big.mean()) < 2e-4
    # known-answer for the hash chain (guards the CUDA generator's restatement in csrc/model.cu)
    v = W.synth(1234, "encoder.conv1.bias", (4,), 0.0, W.B_SCALE)
    assert v.dtype == np.float32 and np.all(np.abs(v) <= 0.02)
