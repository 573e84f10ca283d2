"""log-mel kernel alone (64 x 30 s windows, int16 in): ms per call and GB/s of algorithmic bytes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wdr_b200 as w  # noqa: E402
from oracle import filters  # noqa: E402  (filterbank generator only)

n_mel = int(sys.argv[1]) if len(sys.argv) > 1 else 80
B = 64
fe = w.MelFrontend(filters.whisper_mel_filters(n_mel))
pcm = (torch.randn(B, 480000, device="cuda") * 3000).to(torch.int16)
out = torch.empty(B, n_mel, 3000, device="cuda")
mx = torch.empty(B, device="cuda")
st = torch.cuda.current_stream().cuda_stream
call = lambda: fe.log_mel_batch_dev(pcm.data_ptr(), True, 480000, B, out.data_ptr(), None, mx.data_ptr(), True, st)
for _ in range(5):
    call()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    call()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
alg = B * (480000 * 2 + n_mel * 3000 * 4)
print(f"log-mel {n_mel} bins, {B} windows: {ms * 1e3:.1f} us  {alg / ms / 1e6:.0f} GB/s (incl. the normalise pass)")
fe.close()
