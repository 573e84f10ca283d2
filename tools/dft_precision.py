"""Measures — instead of asserting — what a tensor-core DFT would cost the log-mel in accuracy.

The log-mel kernel is ALU-bound (an fp32 FFT per frame).  The tensor cores could take the 400-point DFT as a GEMM
frames[., 400] x basis[400, 402] (Hann folded into the basis), with the operands split into narrow terms:
  * split-bf16, 3 products (hi*hi + hi*lo + lo*hi): operands carry 16 mantissa bits, products are exact in the fp32 accumulator;
  * 3xTF32 (hi*hi + hi*lo + lo*hi with 11-bit terms): ~21 mantissa bits.
This script emulates both arithmetics bit-faithfully on the CPU (round-to-nearest-even splits, fp32 accumulation), pushes the
result through the same power / mel filterbank / log10 / clamp / (x + 4) / 4 as whisper.cpp and compares against a float64 DFT on the
signals the parity tests use.  Printed: max |error| of the NORMALISED log-mel (the quantity the 1e-4 tolerance is stated on) and the
error as a function of how far a mel bin lies below its chunk's maximum.  No GPU needed:  python tools/dft_precision.py [out.json]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def round_bits(x, bits):
    """fp32 -> nearest value with `bits` explicit mantissa bits (round to nearest even), still stored as fp32."""
    u = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
    drop = 23 - bits
    bias = ((u >> drop) & 1) + (1 << (drop - 1)) - 1
    return (((u + bias) >> drop) << drop).astype(np.uint32).view(np.float32)


def split(x, bits):
    hi = round_bits(x, bits)
    lo = round_bits((x - hi).astype(np.float32), bits)
    return hi, lo


def dft_gemm(frames, basis, bits):
    """3-term split product with fp32 accumulation (each partial product of two `bits`-bit values is exact in fp32 for bits <= 11;
    for bf16, 8 x 8 bits, as well)."""
    fh, fl = split(frames, bits)
    bh, bl = split(basis, bits)
    acc = fh.astype(np.float32) @ bh.astype(np.float32)
    acc = acc + fh @ bl
    acc = acc + fl @ bh
    return acc.astype(np.float32)


def log_mel_from_spec(re, im, filt):
    p = (re.astype(np.float64) ** 2 + im.astype(np.float64) ** 2)
    mel = filt.astype(np.float64) @ p.T
    lm = np.log10(np.maximum(mel, 1e-10))
    lm = np.maximum(lm, lm.max() - 8.0)
    return (lm + 4.0) / 4.0


def main():
    from conftest import synth_audio
    from oracle import filters
    out = {}
    k = np.arange(400)
    hann = 0.5 * (1 - np.cos(2 * np.pi * k / 400))
    ang = 2 * np.pi * np.outer(k, np.arange(201)) / 400
    basis64 = np.concatenate([np.cos(ang) * hann[:, None], -np.sin(ang) * hann[:, None]], 1)  # [400, 402]
    basis32 = basis64.astype(np.float32)
    signals = {"synthetic speech (tests/conftest.py, seed 2000, 30 s)": synth_audio(2000, 30.0).astype(np.float32) / 32768.0,
               "1 kHz full-scale sine": np.sin(2 * np.pi * 1000 * np.arange(160000) / 16000).astype(np.float32),
               "speech at -40 dBFS over a full-scale 200 Hz hum": (synth_audio(2001, 10.0).astype(np.float32) / 32768.0 * 0.01 +
                                                                   0.9 * np.sin(2 * np.pi * 200 * np.arange(160000) / 16000)).astype(np.float32)}
    for n_mel in (80, 128):
        filt = filters.whisper_mel_filters(n_mel)
        for name, x in signals.items():
            pad = np.concatenate([x[200:0:-1], x, np.zeros(200, np.float32)])
            n_fr = (len(pad) - 400) // 160
            idx = np.arange(400)[None, :] + 160 * np.arange(n_fr)[:, None]
            frames = pad[idx].astype(np.float32)
            ref = frames.astype(np.float64) @ basis64
            lm_ref = log_mel_from_spec(ref[:, :201], ref[:, 201:], filt)
            row = {}
            for label, spec in (("fp32 FFT (numpy pocketfft in float32: what the kernel's arithmetic class gives)", None),
                                ("split-bf16 x3 tensor-core DFT", dft_gemm(frames, basis32, 7)),
                                ("3xTF32 tensor-core DFT", dft_gemm(frames, basis32, 10)),
                                ("single bf16 tensor-core DFT", (round_bits(frames, 7) @ round_bits(basis32, 7)).astype(np.float32))):
                if spec is None:
                    f = np.fft.rfft((frames * hann.astype(np.float32)).astype(np.float32), axis=1)
                    f = f.astype(np.complex64)
                    re, im = f.real, f.imag
                else:
                    re, im = spec[:, :201], spec[:, 201:]
                lm = log_mel_from_spec(re, im, filt)
                err = np.abs(lm - lm_ref)
                depth = (lm_ref.max() - lm_ref) * 4.0  # log10 units below the chunk maximum
                bands = {f"{lo}-{lo + 2} below max (log10)": float(err[(depth >= lo) & (depth < lo + 2)].max()) if ((depth >= lo) & (depth < lo + 2)).any() else None
                         for lo in (0, 2, 4, 6)}
                row[label] = {"max_abs_err_normalised_logmel": float(err.max()), "frac_above_1e-4": float((err > 1e-4).mean()), "by_depth": bands}
            out[f"{n_mel} mel | {name}"] = row
    txt = json.dumps(out, indent=1)
    print(txt)
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(txt)


if __name__ == "__main__":
    main()
