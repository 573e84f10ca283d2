"""Kernel micro-timings on one GPU (CUDA events on torch's current stream, which is the stream handed to the
C ABI).  Not the contract bench (bench.py) — a developer tool whose numbers feed DESIGN.md / profiles/."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wdr_b200 as w  # noqa: E402
from oracle import filters  # noqa: E402  (filterbank generator only)


def time_it(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2], ts[0]


def main():
    st = torch.cuda.current_stream().cuda_stream
    res = {}
    for n_mel in (80, 128):
        fe = w.MelFrontend(filters.whisper_mel_filters(n_mel))
        B = 64
        pcm = (torch.randn(B, 480000, device="cuda") * 3000).to(torch.int16)
        pcm_f = pcm.float() / 32768
        out = torch.empty(B, n_mel, 3000, device="cuda")
        mx = torch.empty(B, device="cuda")
        for name, src, is16, bytes_in in (("i16", pcm, True, 2), ("f32", pcm_f, False, 4)):
            med, best = time_it(lambda: fe.log_mel_batch_dev(src.data_ptr(), is16, 480000, B, out.data_ptr(), None, mx.data_ptr(), True, st))
            alg = B * (480000 * bytes_in + n_mel * 3000 * 4)
            res[f"mel{n_mel}_{name}"] = {"ms": med, "ms_best": best, "GBps": alg / med / 1e6, "rtfx": B * 30 / (med / 1e3)}
        fe.close()
    for (M, N, K) in ((96000, 384, 384), (96000, 1536, 384), (96000, 384, 1536), (96000, 1152, 384), (24000, 1280, 1280), (24000, 5120, 1280), (24000, 1280, 5120)):
        A = torch.randn(M, K, device="cuda").bfloat16()
        W = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
        b = torch.randn(N, device="cuda")
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        out32 = torch.randn(M, N, device="cuda", dtype=torch.float32)
        entry = {}
        for name, epi, o, extra in (("bias_bf16", 0, out, None), ("gelu_bf16", 1, out, None), ("resid_f32", 2, out32, out32)):
            med, best = time_it(lambda: w.gemm_bf16_dev(A.data_ptr(), K, M, 1, 0, W.data_ptr(), K, N, K, o.data_ptr(), N, epi, b.data_ptr(),
                                                        None if extra is None else extra.data_ptr(), stream=st))
            entry[name] = {"ms": med, "TFLOPs": 2 * M * N * K / med / 1e9}
        tmed, _ = time_it(lambda: torch.matmul(A, W.T))
        entry["cublas"] = {"ms": tmed, "TFLOPs": 2 * M * N * K / tmed / 1e9}
        res[f"gemm_{M}x{N}x{K}"] = entry
    if len(sys.argv) > 1 and sys.argv[1] == "att":
        pass
    print(json.dumps(res, indent=1))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/microbench.json", "w"), indent=1)


if __name__ == "__main__":
    main()
