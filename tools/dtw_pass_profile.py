"""Two full-pipeline passes; with WDR_PROFILE_DTW_PASS=1 the library brackets the batched DTW pass with cudaProfilerStart/Stop."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wdr_b200 as w
from bench import synth_pcm
arch = sys.argv[1] if len(sys.argv) > 1 else "large-v3"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 120
ctx = w.Context(arch, seed=1234, enable_dtw=True)
st = ctx.create_state()
pcm = torch.from_numpy(synth_pcm(B)).cuda()
p = st.full_params()
for _ in range(2):
    n = st.full_batch_dev(pcm.data_ptr(), B, 480000, p)
torch.cuda.synchronize()
print("segments", n, st.phase_ms())
