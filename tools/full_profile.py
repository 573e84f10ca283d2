"""One full-pipeline pass inside a cudaProfilerStart/Stop range (for `ncu --profile-from-start off`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wdr_b200 as w
from bench import synth_pcm
arch = sys.argv[1] if len(sys.argv) > 1 else "large-v3"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 120
beam = int(sys.argv[3]) if len(sys.argv) > 3 else 0  # > 1: beam search with this width (the crate's default strategy is beam 5)
ctx = w.Context(arch, seed=1234, enable_dtw=True)
st = ctx.create_state()
pcm = torch.from_numpy(synth_pcm(B)).cuda()
p = st.full_params(strategy=1, beam_size=beam) if beam > 1 else st.full_params()
st.full_batch_dev(pcm.data_ptr(), B, 480000, p)
torch.cuda.synchronize()
torch.cuda.profiler.start()
n = st.full_batch_dev(pcm.data_ptr(), B, 480000, p)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("segments", n)
