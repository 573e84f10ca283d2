"""A/B of two builds of the library on the embedding path (measurement helper): tools/_build/old/.../libwdr_b200.so (materialised
im2col) against the in-tree build (implicit GEMM).  Same bf16 products, same k-block order for C >= 64 -> expected bit-identical there;
the C = 32 level groups its taps differently (pairs + a zero phantom tap)."""
import importlib.util, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import synth_audio


def load_pkg(name, so):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "whisper-diarize-rs_b200", "capi.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    m._SO = so
    return m


new = load_pkg("capi_new", os.path.join(ROOT, "whisper-diarize-rs_b200", "csrc", "libwdr_b200.so"))
old = load_pkg("capi_old", os.path.join(ROOT, "tools", "_build", "old", "whisper-diarize-rs_b200", "csrc", "libwdr_b200.so"))
lens = [16000, 300, 48000, 0, 23456, 400, 70000, 160000, 8000, 1234, 5000]
pcm = np.concatenate([synth_audio(20 + i, max(n, 1600) / 16000.0)[:n] for i, n in enumerate(lens)]).astype(np.int16)
off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
out = {}
for tag, m in (("old", old), ("new", new)):
    ex = m.EmbeddingExtractor(seed=1234)
    out[tag], st = ex.compute_batch(pcm, off)
    ex.close()
d = np.abs(out["old"] - out["new"])
print("segments", len(lens), "bit-identical rows", int((d.max(axis=1) == 0).sum()), "max |diff|", float(d.max()), "max |emb|", float(np.abs(out["old"]).max()))
for s in range(len(lens)):
    a, b = out["old"][s], out["new"][s]
    if a.any():
        print(s, lens[s], float(np.abs(a - b).max()), float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b))))
