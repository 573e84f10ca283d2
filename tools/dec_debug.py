import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wdr_b200 as wdr
from oracle import native as oracle, weights as W, vocab as V
arch = sys.argv[1] if len(sys.argv) > 1 else "tiny.en"
w = W.whisper_weights(arch, seed=1234)
a = W.ARCHS[arch]
v = V.special_ids(a["n_vocab"])
rng = np.random.default_rng(11)
B = 2
enc = rng.standard_normal((B, 1500, a["d"])).astype(np.float32)
seqs = np.array([[v["sot"], v["not_"], 1300, 220, 17, v["beg"] + 40], [v["sot"], v["not_"], 5, 6, 7, 8]], np.int32)
aheads = [(l, h) for l in range(a["n_dec"]) for h in (0, 1)]
ctx = wdr.Context(arch, seed=1234, enable_dtw=True)
st = ctx.create_state()
real_aheads = W.ALIGNMENT_HEADS[arch]
logits, ah = st.decode_teacher_forced(seqs, enc=enc, want_logits=True, want_aheads=True, n_aheads=len(real_aheads))
pw = W.pack_decoder(arch, w)
for bf16 in (True, False):
    for b in range(B):
        dec = oracle.Decoder(arch, pw, bf16=bf16)
        dec.set_audio(enc[b])
        for i, t in enumerate(seqs[b]):
            lg, pr = dec.step(int(t), i, aheads=real_aheads)
            e = np.abs(lg - logits[b, i])
            print(f"bf16={bf16} b={b} i={i} logit err max {e.max():.2e} mean {e.mean():.2e} scale {np.abs(lg).max():.3f}  "
                  + " ".join(f"L{l}h{h}:{np.abs(pr[k]-ah[b,k,i]).max():.1e}/{pr[k].max():.1e}" for k, (l, h) in enumerate(real_aheads)))
