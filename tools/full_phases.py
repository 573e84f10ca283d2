"""Per-step phase times of the full pipeline (wdr_full_get_phase_ms) + host wall time, several steps in a row."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import wdr_b200 as w
from bench import synth_pcm

arch = sys.argv[1] if len(sys.argv) > 1 else "large-v3"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 120
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
ctx = w.Context(arch, seed=1234, enable_dtw=True)
st = ctx.create_state()
pcm = torch.from_numpy(synth_pcm(B)).cuda()
p = st.full_params()
for i in range(steps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = st.full_batch_dev(pcm.data_ptr(), B, 480000, p)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) * 1e3
    ph = st.phase_ms()
    print(f"step {i}: wall {dt:.0f} ms, segments {n}, phases {{" + ", ".join(f"{k}: {v:.1f}" if k != "decode_steps" else f"{k}: {v}" for k, v in ph.items()) + "}", flush=True)
