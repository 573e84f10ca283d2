"""CPU suite: the C-ABI library loads, exports every symbol include/wdr.h declares, and fails loudly
(WDR_ERR_NO_DEVICE) instead of falling back when there is no GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    syms = set()
    inc = os.path.join(ROOT, "include")
    for name in os.listdir(inc):
        if name.endswith(".h"):
            text = open(os.path.join(inc, name)).read()
            text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
            syms |= set(re.findall(r"\b(wdr_[a-z0-9_]+)\s*\(", text))
    return sorted(syms)


def test_library_exports_every_declared_symbol(wdr):
    lib = ctypes.CDLL(wdr.lib_path())
    syms = declared_symbols()
    assert len(syms) >= 15
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, f"declared in include/*.h but not exported: {missing}"


def test_version_and_shape_helpers(wdr):
    assert "sm_100a" in wdr.version()
    assert wdr.mel_n_len(480000) == 6000 and wdr.mel_n_len(0) == 3000
    assert wdr.fbank_frames(399) == 0 and wdr.fbank_frames(400) == 1 and wdr.fbank_frames(16000) == 98


def test_no_cpu_fallback(wdr):
    if wdr.device_count() > 0:
        pytest.skip("a GPU is present; the no-device path is exercised on the CPU box")
    with pytest.raises(wdr.WdrError) as e:
        wdr.median_filter(np.zeros((1, 2, 16), np.float32))
    assert e.value.code == -2
    with pytest.raises(wdr.WdrError):
        wdr.MelFrontend(np.zeros((80, 201), np.float32))
    with pytest.raises(wdr.WdrError):
        wdr.dtw(np.zeros((3, 4), np.float32))


def test_product_never_touches_oracle():
    """The product path must not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "whisper-diarize-rs_b200")
    bad = []
    for d in (pkg, os.path.join(ROOT, "host"), os.path.join(ROOT, "include")):
        for base, _, files in os.walk(d):
            if "build" in base:
                continue
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp", "Makefile")):
                    txt = open(os.path.join(base, f), errors="replace").read()
                    if re.search(r"(from|import)\s+oracle|oracle/|liboracle|oracle_[a-z_]+\(", txt):
                        bad.append(os.path.join(base, f))
    assert not bad, bad
