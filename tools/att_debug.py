"""Runs the attention kernel on one (B,T,H) config given on the command line; prints max error."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wdr_b200 as wdr
B, T, H = map(int, sys.argv[1:4])
g = torch.Generator(device="cuda").manual_seed(1)
d = 64 * H; M = B * T; ldt = (M + 7) // 8 * 8
qkv = torch.randn(M, 3 * d, device="cuda", generator=g) * 1.5
qk = qkv[:, :2 * d].contiguous().bfloat16()
vt = torch.zeros(d, ldt, device="cuda", dtype=torch.bfloat16); vt[:, :M] = qkv[:, 2 * d:].T.bfloat16()
out = torch.full((M, d), float("nan"), device="cuda", dtype=torch.bfloat16)
wdr.encoder_attention_dev(qk.data_ptr(), vt.data_ptr(), ldt, B, T, H, d, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
q = qk[:, :d].float().view(B, T, H, 64).transpose(1, 2); k = qk[:, d:].float().view(B, T, H, 64).transpose(1, 2)
v = vt[:, :M].float().T.reshape(B, T, H, 64).transpose(1, 2)
ref = torch.nn.functional.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(M, d)
print("OK", B, T, H, "maxerr", (out.float() - ref).abs().max().item(), "refmax", ref.abs().max().item())
