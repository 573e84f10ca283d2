"""Writes whisper.cpp `ggml-<model>.bin` checkpoints from a dict of OpenAI-named fp32 tensors (test infrastructure): the same
container `convert-pt-to-ggml.py` produces — magic, 11 hyper-parameters, mel filterbank, vocabulary, then per tensor
(n_dims, name_len, type, ne[] innermost first, name, data).  f16 files keep 1-D tensors, the conv biases and the positional
embeddings in f32, as the converter does."""
import struct

import numpy as np

F32_ALWAYS = ("encoder.conv1.bias", "encoder.conv2.bias", "encoder.positional_embedding", "decoder.positional_embedding")


def write_ggml(path, arch, weights, filters, tokens, use_f16=False):
    """arch: dict(d, n_head, n_enc, n_dec, n_mel, n_vocab) (oracle.weights.ARCHS entry); weights: name -> fp32 array (OpenAI layout);
    filters [n_mel, 201] fp32; tokens: list of bytes (the text vocabulary, ids 0 .. len-1)."""
    with open(path, "wb") as f:
        f.write(struct.pack("<I", 0x67676D6C))
        f.write(struct.pack("<11i", arch["n_vocab"], 1500, arch["d"], arch["n_head"], arch["n_enc"], 448, arch["d"], arch["n_head"], arch["n_dec"],
                            arch["n_mel"], 1 if use_f16 else 0))
        filt = np.ascontiguousarray(filters, np.float32)
        f.write(struct.pack("<2i", filt.shape[0], filt.shape[1]))
        f.write(filt.tobytes())
        f.write(struct.pack("<i", len(tokens)))
        for t in tokens:
            f.write(struct.pack("<I", len(t)))
            f.write(t)
        for name, w in weights.items():
            a = np.ascontiguousarray(w, np.float32)
            if name in ("encoder.conv1.bias", "encoder.conv2.bias"):
                a = a.reshape(-1, 1)  # the converter stores the conv biases as [n_state, 1]
            as_f16 = use_f16 and a.ndim >= 2 and name not in F32_ALWAYS
            nb = name.encode()
            f.write(struct.pack("<3i", a.ndim, len(nb), 1 if as_f16 else 0))
            for dim in reversed(a.shape):
                f.write(struct.pack("<i", dim))
            f.write(nb)
            f.write((a.astype(np.float16) if as_f16 else a).tobytes())
