import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, time
import wdr_b200 as w
from bench import synth_pcm
arch = sys.argv[1] if len(sys.argv) > 1 else "large-v3"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 120
ctx = w.Context(arch, seed=1234, enable_dtw=True)
st = ctx.create_state()
pcm = torch.from_numpy(synth_pcm(B)).cuda()
p = st.full_params()
st.full_batch_dev(pcm.data_ptr(), B, 480000, p)
t = time.time(); st.full_batch_dev(pcm.data_ptr(), B, 480000, p); print("wall", time.time() - t)
info = [st.chunk_info(i) for i in range(B)]
ns = np.array([c["n_sampled"] for c in info])
print("n_sampled: mean", ns.mean(), "min", ns.min(), "max", ns.max(), "hist", np.histogram(ns, bins=[0, 20, 40, 80, 120, 160, 200, 221])[0])
print("failed", sum(c["failed"] for c in info), "completed", sum(c["completed"] for c in info))
segs = st.segments()
nt = np.array([sum(1 for t in s["tokens"] if t.id < 50257) for s in segs])
print("text tokens/window: mean", nt.mean(), "max", nt.max())
st.profile_enable(True); st.profile_collect()
st.full_batch_dev(pcm.data_ptr(), B, 480000, p)
pr = st.profile_collect()
print({k: (round(v["ms"], 1), v["records"]) for k, v in pr.items()})
