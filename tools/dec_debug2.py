import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["WDR_DEBUG_DEC_LAYERS"] = "1"
import numpy as np
import wdr_b200 as wdr
from oracle import weights as W, vocab as V
arch = "tiny.en"
w = W.whisper_weights(arch, seed=1234)
a = W.ARCHS[arch]; d = a["d"]; H = a["n_head"]
v = V.special_ids(a["n_vocab"])
rng = np.random.default_rng(11)
B = 2
enc = rng.standard_normal((B, 1500, d)).astype(np.float32)
seqs = np.array([[v["sot"]], [1300]], np.int32)
ctx = wdr.Context(arch, seed=1234, enable_dtw=True)
st = ctx.create_state()
logits, _ = st.decode_teacher_forced(seqs, enc=enc, want_logits=True)
L = wdr.load()
L.wdr_debug_decoder_read.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float), C.c_int64]
def rd(which, n):
    out = np.empty(n, np.float32)
    assert L.wdr_debug_decoder_read(st._h, which, out.ctypes.data_as(C.POINTER(C.c_float)), n) == 0
    return out
bf = W.bf16_round
def ln(x, g, b):
    m = x.mean(dtype=np.float64); var = ((x - np.float32(m)) ** 2).mean(dtype=np.float64)
    return ((x - np.float32(m)) / np.sqrt(np.float32(var) + np.float32(1e-5)) * g + b).astype(np.float32)
def gelu(x):
    return (0.5 * x * (1 + np.tanh(0.7978845608028654 * x * (1 + 0.044715 * x * x)))).astype(np.float32)
P = "decoder.blocks.0."
ckv = rd(4, B * 1500 * 2 * d).reshape(B, 1500, 2 * d)
for b in range(B):
    e = bf(enc[b])
    Kc = bf(e @ w[P + "cross_attn.key.weight"].T)
    Vc = bf(e @ w[P + "cross_attn.value.weight"].T + w[P + "cross_attn.value.bias"])
    print("b", b, "Kc err", np.abs(Kc - ckv[b, :, :d]).max(), "mismatch frac", (Kc != ckv[b, :, :d]).mean(), "Vc err", np.abs(Vc - ckv[b, :, d:]).max(), (Vc != ckv[b, :, d:]).mean())
    tok = int(seqs[b, 0])
    x = w["decoder.token_embedding.weight"][tok] + w["decoder.positional_embedding"][0]
    h = bf(ln(x, w[P + "attn_ln.weight"], w[P + "attn_ln.bias"]))
    vv = bf(h @ w[P + "attn.value.weight"].T + w[P + "attn.value.bias"])
    att = vv  # single position
    x = x + att @ w[P + "attn.out.weight"].T + w[P + "attn.out.bias"]
    h = bf(ln(x, w[P + "cross_attn_ln.weight"], w[P + "cross_attn_ln.bias"]))
    q = h @ w[P + "cross_attn.query.weight"].T + w[P + "cross_attn.query.bias"]
    att = np.zeros(d, np.float32)
    for hh in range(H):
        s = (Kc[:, hh * 64:(hh + 1) * 64] @ q[hh * 64:(hh + 1) * 64]) * np.float32(0.125)
        p = np.exp(s - s.max()); p = p / p.sum()
        att[hh * 64:(hh + 1) * 64] = p @ Vc[:, hh * 64:(hh + 1) * 64]
    att = bf(att)
    g_att = rd(2, B * d).reshape(B, d)[b]
    print("  cross att err", np.abs(att - g_att).max(), "scale", np.abs(att).max(), "mismatch", (att != g_att).mean())
    x = x + att @ w[P + "cross_attn.out.weight"].T + w[P + "cross_attn.out.bias"]
    h = bf(ln(x, w[P + "mlp_ln.weight"], w[P + "mlp_ln.bias"]))
    ff = bf(gelu(h @ w[P + "mlp.0.weight"].T + w[P + "mlp.0.bias"]))
    g_ff = rd(3, B * 4 * d).reshape(B, 4 * d)[b]
    print("  ff err", np.abs(ff - g_ff).max(), "scale", np.abs(ff).max(), "mismatch", (ff != g_ff).mean())
    x = x + ff @ w[P + "mlp.2.weight"].T + w[P + "mlp.2.bias"]
    g_x = rd(0, B * d).reshape(B, d)[b]
    print("  x err", np.abs(x - g_x).max(), "scale", np.abs(x).max())
    h = bf(ln(x, w["decoder.ln.weight"], w["decoder.ln.bias"]))
    g_h = rd(1, B * d).reshape(B, d)[b]
    print("  h err", np.abs(h - g_h).max(), "mismatch", (h != g_h).mean())
    lg = h @ w["decoder.token_embedding.weight"].T
    print("  logits err", np.abs(lg - logits[b, 0]).max(), "mean", np.abs(lg - logits[b, 0]).mean(), "scale", np.abs(lg).max())
    lg2 = g_h @ w["decoder.token_embedding.weight"].T
    print("  logits err using GPU h", np.abs(lg2 - logits[b, 0]).max(), "mean", np.abs(lg2 - logits[b, 0]).mean())
