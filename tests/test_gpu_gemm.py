"""GPU tests of the tcgen05 bf16 GEMM (csrc/gemm.cu) through the C ABI, against a plain PyTorch fp32 reference of
the same op on the same bf16-rounded operands.  Tolerance: fp32 accumulation of bf16 products is exact up to
summation order, so outputs agree to ~1e-5 relative before the final bf16 rounding (2^-8 relative) — the bf16
outputs are compared with rtol 1e-2 (north_star: encoder within 1e-2 relative), fp32 outputs with 1e-4."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def gelu_tanh(x):
    return 0.5 * x * (1 + torch.tanh(0.7978845608028654 * x * (1 + 0.044715 * x * x)))


def run(wdr, A, W, epilogue=0, bias=None, extra=None, rows_per_batch=None, n_batch=1, n_split=0, out_dtype=torch.bfloat16,
        kb_per_tap=0, a_cols=0, K=None, lda=None, a_batch_stride=0, M_out=None):
    N = W.shape[0]
    K = K or A.shape[-1]
    rows_per_batch = rows_per_batch or A.shape[0]
    M = M_out or rows_per_batch * n_batch
    out = torch.full((M, N), float("nan"), dtype=out_dtype, device="cuda")
    out_t = torch.full((max(N - n_split, 1), M), float("nan"), dtype=torch.bfloat16, device="cuda") if epilogue == 4 else None
    wdr.gemm_bf16_dev(A.data_ptr(), lda or A.stride(-2), rows_per_batch, n_batch, a_batch_stride, W.data_ptr(), W.stride(0), N, K,
                      out.data_ptr(), N, epilogue, None if bias is None else bias.data_ptr(), None if extra is None else extra.data_ptr(),
                      None if out_t is None else out_t.data_ptr(), M, n_split, kb_per_tap, a_cols, 0, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return out, out_t


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 128, 128), (1000, 384, 384), (3000, 1536, 384), (257, 200, 240), (96000, 384, 384)])
def test_gemm_bias(wdr, M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
    b = torch.randn(N, device="cuda", generator=g)
    out, _ = run(wdr, A, W, 0, b)
    ref = A.float() @ W.float().T + b
    assert torch.isfinite(out.float()).all()
    err = (out.float() - ref).abs().max().item()
    assert err <= 1e-2 * ref.abs().max().item() + 1e-3, err
    out32, _ = run(wdr, A, W, 5, b, out_dtype=torch.float32)
    assert (out32 - ref).abs().max().item() <= 1e-4 * ref.abs().max().item() + 1e-5


def test_gemm_batched_rows_gelu_and_residual(wdr):
    g = torch.Generator(device="cuda").manual_seed(7)
    B, R, K, N = 3, 1500, 384, 512
    A = (torch.randn(B * R, K, device="cuda", generator=g)).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
    b = torch.randn(N, device="cuda", generator=g) * 0.1
    out, _ = run(wdr, A, W, 1, b, rows_per_batch=R, n_batch=B, a_batch_stride=R * K)
    ref = gelu_tanh(A.float() @ W.float().T + b)
    assert (out.float() - ref).abs().max().item() <= 1e-2 * ref.abs().max().item() + 1e-3
    resid = torch.randn(B * R, N, device="cuda", generator=g)
    out32, _ = run(wdr, A, W, 2, b, extra=resid, rows_per_batch=R, n_batch=B, a_batch_stride=R * K, out_dtype=torch.float32)
    ref2 = resid + A.float() @ W.float().T + b
    assert (out32 - ref2).abs().max().item() <= 1e-4 * ref2.abs().max().item() + 1e-5
    pos = torch.randn(R, N, device="cuda", generator=g)
    out3, _ = run(wdr, A, W, 3, b, extra=pos, rows_per_batch=R, n_batch=B, a_batch_stride=R * K, out_dtype=torch.float32)
    ref3 = ref.view(B, R, N) + pos
    assert (out3.view(B, R, N) - ref3).abs().max().item() <= 1e-4 * ref3.abs().max().item() + 1e-4


def test_gemm_qkv_transposed_split(wdr):
    g = torch.Generator(device="cuda").manual_seed(9)
    M, K, d = 1500 * 2, 384, 384
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    W = (torch.randn(3 * d, K, device="cuda", generator=g) * 0.05).bfloat16()
    b = torch.randn(3 * d, device="cuda", generator=g) * 0.1
    out, out_t = run(wdr, A, W, 4, b, n_split=2 * d)
    ref = A.float() @ W.float().T + b
    tol = 1e-2 * ref.abs().max().item() + 1e-3
    assert (out[:, : 2 * d].float() - ref[:, : 2 * d]).abs().max().item() <= tol
    assert (out_t.float() - ref[:, 2 * d:].T).abs().max().item() <= tol


def test_gemm_conv_tap_mode(wdr):
    """conv1d(k=3, stride=2, pad=1) as implicit GEMM over pair-rows (encoder conv2)."""
    g = torch.Generator(device="cuda").manual_seed(11)
    B, C, T, N = 2, 384, 3000, 384
    x = torch.randn(B, T, C, device="cuda", generator=g).bfloat16()                   # frames, token-major
    w = (torch.randn(N, C, 3, device="cuda", generator=g) * 0.03).bfloat16()          # [out][in][tap]
    b = torch.randn(N, device="cuda", generator=g) * 0.1
    P = torch.zeros(B, T + 2, C, device="cuda", dtype=torch.bfloat16)                  # row 0 = left pad, row T+1 = spare
    P[:, 1: T + 1] = x
    Wk = w.permute(0, 2, 1).contiguous().view(N, 3 * C)                                # K index = tap*C + c
    rows = T // 2
    out, _ = run(wdr, P.view(B * (T + 2) // 2, 2 * C), Wk, 1, b, rows_per_batch=rows, n_batch=B, a_batch_stride=(T + 2) * C,
                 K=3 * C, lda=2 * C, kb_per_tap=2 * C // 64, a_cols=2 * C, M_out=B * rows)
    ref = torch.nn.functional.conv1d(x.float().transpose(1, 2), w.float(), b, stride=2, padding=1).transpose(1, 2)
    ref = gelu_tanh(ref).reshape(B * rows, N)
    assert (out.float() - ref).abs().max().item() <= 1e-2 * ref.abs().max().item() + 1e-3
