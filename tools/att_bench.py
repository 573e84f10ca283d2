"""Encoder attention kernel alone (large-v3 shape: 20 heads, T = 1500) on B windows: ms per call and TFLOP/s (4 T^2 d per window)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wdr_b200 as w  # noqa: E402

B, T, H = int(sys.argv[1]) if len(sys.argv) > 1 else 60, 1500, 20
d = H * 64
T_pad = (T + 7) // 8 * 8
ldt = B * T_pad
qk = (torch.randn(B * T, 2 * d, device="cuda") * 0.5).bfloat16()
vt = torch.randn(d, ldt, device="cuda").bfloat16()
out = torch.empty(B * T, d, device="cuda", dtype=torch.bfloat16)
st = torch.cuda.current_stream().cuda_stream
call = lambda: w.encoder_attention_dev(qk.data_ptr(), vt.data_ptr(), ldt, B, T, H, d, out.data_ptr(), st)
for _ in range(3):
    call()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    call()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"attention B={B}: {ms:.3f} ms  {4 * T * T * d * B / ms / 1e9:.1f} TFLOP/s")
