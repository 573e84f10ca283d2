"""Seeded synthetic Whisper weights (test infrastructure).

No checkpoint can exist in this environment (no network), so both the CUDA library and the oracle synthesise the
weights of a named architecture from (seed, tensor name, flat index) with the same counter-based generator:
    key  = splitmix64(fnv1a64(name) ^ splitmix64(seed))
    u    = (int(splitmix64(key + i) >> 40) - 2^23) / 2^23          in [-1, 1), exact in fp32
    w[i] = offset + u * scale                                       one fp32 multiply-add
Matrices that feed tensor-core GEMMs are rounded to bf16 once (round-to-nearest-even): the synthetic "checkpoint"
is bf16, as a real one would be f16 for whisper.cpp.  Tensor names and layouts are OpenAI Whisper's.
"""
import numpy as np

M64 = np.uint64(0xFFFFFFFFFFFFFFFF)

ARCHS = {
    "tiny.en": dict(d=384, n_head=6, n_enc=4, n_dec=4, n_mel=80, n_vocab=51864),
    "tiny": dict(d=384, n_head=6, n_enc=4, n_dec=4, n_mel=80, n_vocab=51865),
    "base.en": dict(d=512, n_head=8, n_enc=6, n_dec=6, n_mel=80, n_vocab=51864),
    "base": dict(d=512, n_head=8, n_enc=6, n_dec=6, n_mel=80, n_vocab=51865),
    "small.en": dict(d=768, n_head=12, n_enc=12, n_dec=12, n_mel=80, n_vocab=51864),
    "small": dict(d=768, n_head=12, n_enc=12, n_dec=12, n_mel=80, n_vocab=51865),
    "medium.en": dict(d=1024, n_head=16, n_enc=24, n_dec=24, n_mel=80, n_vocab=51864),
    "medium": dict(d=1024, n_head=16, n_enc=24, n_dec=24, n_mel=80, n_vocab=51865),
    "large-v3": dict(d=1280, n_head=20, n_enc=32, n_dec=32, n_mel=128, n_vocab=51866),
    "large-v3-turbo": dict(d=1280, n_head=20, n_enc=32, n_dec=4, n_mel=128, n_vocab=51866),
}

W_SCALE = np.float32(0.034641016151377546)
B_SCALE = np.float32(0.02)


def _splitmix64(x):
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & M64
        z = x
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & M64
        return z ^ (z >> np.uint64(31))


def _fnv1a(name):
    h = 0xCBF29CE484222325
    for ch in name.encode():
        h ^= ch
        h = (h * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return np.uint64(h)


def bf16_round(x):
    u = np.ascontiguousarray(x, np.float32).view(np.uint32)
    r = ((u >> np.uint32(16)) & np.uint32(1)) + np.uint32(0x7FFF)
    return ((u + r) & np.uint32(0xFFFF0000)).view(np.float32)


def _synth_native(key, n, offset, scale, bf16):
    """C fast path (oracle_synth_fill, same arithmetic); None if the oracle library is not built."""
    try:
        import ctypes as C
        from . import native
        L = native.lib()
        fn = L.oracle_synth_fill
        fn.argtypes = [C.POINTER(C.c_float), C.c_int64, C.c_uint64, C.c_float, C.c_float, C.c_int]
        fn.restype = None
        out = np.empty(n, np.float32)
        fn(out.ctypes.data_as(C.POINTER(C.c_float)), n, int(key), float(np.float32(offset)), float(np.float32(scale)), int(bf16))
        return out
    except Exception:
        return None


def synth(seed, name, shape, offset=0.0, scale=W_SCALE, bf16=False, native_ok=True):
    key = _splitmix64(_fnv1a(name) ^ _splitmix64(np.uint64(seed)))
    n = int(np.prod(shape))
    if native_ok and n >= 65536:
        v = _synth_native(key, n, offset, scale, bf16)
        if v is not None:
            return v.reshape(shape)
    with np.errstate(over="ignore"):
        z = _splitmix64((key + np.arange(n, dtype=np.uint64)) & M64)
    k = (z >> np.uint64(40)).astype(np.int64) - 8388608
    u = k.astype(np.float32) * np.float32(1.0 / 8388608.0)
    v = (np.float32(offset) + u * np.float32(scale)).astype(np.float32)
    if bf16:
        v = bf16_round(v)
    return v.reshape(shape)


def sinusoids(length, channels):
    half = channels // 2
    inc = np.float32(np.log(np.float32(10000.0))) / np.float32(half - 1)
    inv = np.exp(-inc * np.arange(half, dtype=np.float32)).astype(np.float32)
    st = np.arange(length, dtype=np.float32)[:, None] * inv[None, :]
    return np.concatenate([np.sin(st), np.cos(st)], axis=1).astype(np.float32)


def whisper_weights(arch, seed=1234):
    """dict name -> fp32 array in OpenAI Whisper layout (matrices hold bf16-representable values)."""
    a = ARCHS[arch]
    d = a["d"]
    w = {}

    def mat(name, shape, scale=W_SCALE):
        w[name] = synth(seed, name, shape, 0.0, scale, bf16=True)

    def vec(name, n, offset=0.0, scale=B_SCALE):
        w[name] = synth(seed, name, (n,), offset, scale)

    def block(prefix, n_layer, cross):
        out_scale = np.float32(W_SCALE / np.sqrt(np.float32(2.0 * n_layer)))
        for l in range(n_layer):
            p = f"{prefix}.blocks.{l}."
            vec(p + "attn_ln.weight", d, 1.0, 0.1)
            vec(p + "attn_ln.bias", d, 0.0, 0.1)
            mat(p + "attn.query.weight", (d, d))
            vec(p + "attn.query.bias", d)
            mat(p + "attn.key.weight", (d, d))
            mat(p + "attn.value.weight", (d, d))
            vec(p + "attn.value.bias", d)
            mat(p + "attn.out.weight", (d, d), out_scale)
            vec(p + "attn.out.bias", d)
            if cross:
                vec(p + "cross_attn_ln.weight", d, 1.0, 0.1)
                vec(p + "cross_attn_ln.bias", d, 0.0, 0.1)
                mat(p + "cross_attn.query.weight", (d, d))
                vec(p + "cross_attn.query.bias", d)
                mat(p + "cross_attn.key.weight", (d, d))
                mat(p + "cross_attn.value.weight", (d, d))
                vec(p + "cross_attn.value.bias", d)
                mat(p + "cross_attn.out.weight", (d, d), out_scale)
                vec(p + "cross_attn.out.bias", d)
            vec(p + "mlp_ln.weight", d, 1.0, 0.1)
            vec(p + "mlp_ln.bias", d, 0.0, 0.1)
            mat(p + "mlp.0.weight", (4 * d, d))
            vec(p + "mlp.0.bias", 4 * d)
            mat(p + "mlp.2.weight", (d, 4 * d), out_scale)
            vec(p + "mlp.2.bias", d)

    mat("encoder.conv1.weight", (d, a["n_mel"], 3))
    vec("encoder.conv1.bias", d)
    mat("encoder.conv2.weight", (d, d, 3))
    vec("encoder.conv2.bias", d)
    w["encoder.positional_embedding"] = sinusoids(1500, d)
    block("encoder", a["n_enc"], cross=False)
    vec("encoder.ln_post.weight", d, 1.0, 0.1)
    vec("encoder.ln_post.bias", d, 0.0, 0.1)
    mat("decoder.token_embedding.weight", (a["n_vocab"], d))
    w["decoder.positional_embedding"] = synth(seed, "decoder.positional_embedding", (448, d), 0.0, 0.017320508)
    block("decoder", a["n_dec"], cross=True)
    vec("decoder.ln.weight", d, 1.0, 0.1)
    vec("decoder.ln.bias", d, 0.0, 0.1)
    return w


def pack_encoder(arch, w):
    """Flat fp32 array in the order oracle_whisper_encode walks."""
    a = ARCHS[arch]
    parts = [w["encoder.conv1.weight"], w["encoder.conv1.bias"], w["encoder.conv2.weight"], w["encoder.conv2.bias"],
             w["encoder.positional_embedding"]]
    for l in range(a["n_enc"]):
        p = f"encoder.blocks.{l}."
        for n in ("attn_ln.weight", "attn_ln.bias", "attn.query.weight", "attn.query.bias", "attn.key.weight", "attn.value.weight",
                  "attn.value.bias", "attn.out.weight", "attn.out.bias", "mlp_ln.weight", "mlp_ln.bias", "mlp.0.weight", "mlp.0.bias",
                  "mlp.2.weight", "mlp.2.bias"):
            parts.append(w[p + n])
    parts += [w["encoder.ln_post.weight"], w["encoder.ln_post.bias"]]
    return np.concatenate([np.ravel(x) for x in parts]).astype(np.float32)


def pack_decoder(arch, w):
    """Flat fp32 array in the order oracle_dec_create walks (oracle/wdr_oracle_full.c)."""
    a = ARCHS[arch]
    parts = [w["decoder.token_embedding.weight"], w["decoder.positional_embedding"]]
    for l in range(a["n_dec"]):
        p = f"decoder.blocks.{l}."
        for n in ("attn_ln.weight", "attn_ln.bias", "attn.query.weight", "attn.query.bias", "attn.key.weight", "attn.value.weight",
                  "attn.value.bias", "attn.out.weight", "attn.out.bias", "cross_attn_ln.weight", "cross_attn_ln.bias",
                  "cross_attn.query.weight", "cross_attn.query.bias", "cross_attn.key.weight", "cross_attn.value.weight",
                  "cross_attn.value.bias", "cross_attn.out.weight", "cross_attn.out.bias", "mlp_ln.weight", "mlp_ln.bias",
                  "mlp.0.weight", "mlp.0.bias", "mlp.2.weight", "mlp.2.bias"):
            parts.append(w[p + n])
    parts += [w["decoder.ln.weight"], w["decoder.ln.bias"]]
    return np.concatenate([np.ravel(x) for x in parts]).astype(np.float32)


# (layer, head) alignment heads per DTW preset (OpenAI _ALIGNMENT_HEADS)
ALIGNMENT_HEADS = {
    "tiny.en": [(1, 0), (2, 0), (2, 5), (3, 0), (3, 1), (3, 2), (3, 3), (3, 4)],
    "tiny": [(2, 2), (3, 0), (3, 2), (3, 3), (3, 4), (3, 5)],
    "base.en": [(3, 3), (4, 7), (5, 1), (5, 5), (5, 7)],
    "base": [(3, 1), (4, 2), (4, 3), (4, 7), (5, 1), (5, 2), (5, 4), (5, 6)],
    "small.en": [(6, 6), (7, 0), (7, 3), (7, 8), (8, 2), (8, 5), (8, 7), (9, 0), (9, 4), (9, 8), (9, 10), (10, 0), (10, 1), (10, 2),
                 (10, 3), (10, 6), (10, 11), (11, 2), (11, 4)],
    "small": [(5, 3), (5, 9), (8, 0), (8, 4), (8, 7), (8, 8), (9, 0), (9, 7), (9, 9), (10, 5)],
    "medium.en": [(11, 4), (14, 1), (14, 12), (14, 14), (15, 4), (16, 0), (16, 4), (16, 9), (17, 12), (17, 14), (18, 7), (18, 10),
                  (18, 15), (20, 0), (20, 3), (20, 9), (20, 14), (21, 12)],
    "medium": [(13, 15), (15, 4), (15, 15), (16, 1), (20, 0), (23, 4)],
    "large-v3": [(7, 0), (10, 17), (12, 18), (13, 12), (16, 1), (17, 14), (19, 11), (21, 4), (24, 1), (25, 6)],
    "large-v3-turbo": [(2, 4), (2, 11), (3, 3), (3, 6), (3, 11), (3, 14)],
}
