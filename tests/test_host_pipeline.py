"""CPU known-answer tests for the host-side mirror of the reference's own Rust around the boundary:
src/utils.rs:3-59 and src/transcribe.rs:171-320, 397-459 (hostmirror/host.py)."""
from types import SimpleNamespace as NS

import importlib

import numpy as np

H = importlib.import_module("hostmirror.host")


def test_calculate_dtw_mem_size_table():
    MB = 1024 * 1024
    # 960 000 samples -> 6000 frames -> 24 MB + 6000*96*16 B + 24 kB = 33.2 MB -> 40 MB after 8 MB alignment
    assert H.calculate_dtw_mem_size(960_000) == 40 * MB
    assert H.calculate_dtw_mem_size(0) == 24 * MB
    assert H.calculate_dtw_mem_size(1) == 32 * MB            # 24 MB + one frame: aligned UP
    assert H.calculate_dtw_mem_size(16000 * 3600 * 8) == 768 * MB  # ceiling (already 8 MB aligned)
    # band switches at 15 000 / 45 000 frames
    a, b = H.calculate_dtw_mem_size(15_000 * 160), H.calculate_dtw_mem_size(15_001 * 160)
    assert a == 48 * MB and b == 56 * MB


def test_control_tokens():
    for s in ("[_BEG_]", "[_TT_320]", " [_EOT_] ", "[_LANG_EN]", "[_extra_token_50360]".upper()):
        assert H.is_whole_control_token(s), s
    for s in ("[_beg_]", "[BEG]", "[_]", "hello", "[_TT_320] x", "[_extra_token_50360]"):
        assert not H.is_whole_control_token(s), s
    assert H.strip_embedded_control_markers("he[_TT_12]llo [_x] [_BEG_]!") == "hello [_x] !"
    assert H.strip_embedded_control_markers("[_unclosed") == "[_unclosed"


def _td(p, t0, t1, t_dtw):
    return NS(p=p, t0=t0, t1=t1, t_dtw=t_dtw)


def test_get_token_timestamps_anchor_rules():
    texts = ["[_BEG_]", " Hello", " wor", "ld", "[_TT_150]", " \0"]
    data = [_td(0.9, 0, 0, -1), _td(0.5, 10, 50, 20), _td(0.6, 50, 90, 60), _td(0.7, 90, 120, -1), _td(0.8, 120, 120, 300), _td(0.1, 0, 0, 5)]
    w = H.get_token_timestamps(texts, data)
    assert [x["text"] for x in w] == [" Hello", " wor", "ld"]
    # first token: no previous anchor -> t0; end = midpoint(0.20, 0.60)
    assert w[0]["start"] == 0.1 and abs(w[0]["end"] - 0.4) < 1e-12
    # middle: start = midpoint with previous; next has no anchor -> t1
    assert abs(w[1]["start"] - 0.4) < 1e-12 and w[1]["end"] == 0.9
    # no anchor here -> t0 / t1
    assert w[2]["start"] == 0.9 and w[2]["end"] == 1.2
    assert [x["probability"] for x in w] == [0.5, 0.6, 0.7]
    assert H.get_token_timestamps(["[_BEG_]"], [_td(1, 0, 0, -1)]) == []


def test_interpolate_word_timestamps():
    w = H.interpolate_word_timestamps("ab  cdef , x", 10.0, 14.0)  # weights 2, 4, 1 (punctuation -> 1), 1
    assert [x["text"] for x in w] == ["ab", "cdef", ",", "x"]
    assert [x["start"] for x in w] == [10.0, 11.0, 13.0, 13.5] and w[-1]["end"] == 14.0 and w[0]["end"] == 11.0
    assert H.interpolate_word_timestamps("a b", 3.0, 3.0) == [] and H.interpolate_word_timestamps("   ", 0.0, 1.0) == []


def test_assemble_segments_offsets_and_overlap_clipping():
    segs = [dict(text=" Hi there", t0=0, t1=100, token_text=[" Hi", " there"], tokens=[_td(0.9, 0, 40, -1), _td(0.8, 40, 100, -1)])]
    out = H.assemble_segments(segs, 5.0, [])
    assert out[0]["text"] == "Hi there" and out[0]["start"] == 5.0 and out[0]["end"] == 6.0
    nxt = [dict(text=" ok", t0=0, t1=50, token_text=[" ok"], tokens=[_td(0.7, 0, 50, -1)])]
    out = H.assemble_segments(nxt, 5.8, out)  # starts before the previous segment ended: previous end and its last word are clipped
    assert out[0]["end"] == 5.8 and out[0]["words"][-1]["end"] == 5.8 and out[1]["start"] == 5.8
    # a segment whose tokens are all control markers keeps the approximate bounds and has no words
    ctl = [dict(text="", t0=10, t1=20, token_text=["[_BEG_]"], tokens=[_td(1.0, 10, 20, -1)])]
    out = H.assemble_segments(ctl, 0.0, out)
    assert out[-1]["words"] is None and out[-1]["start"] == 0.1 and out[-1]["end"] == 0.2


def test_vad_slice_indices_round_like_rust_above_2_pow_23():
    """src/vad.rs:69-70: ((t as f32 * SR).round()).clamp(0, n) as usize.  Above 2^23 every f32 is an integer, so round() is the
    identity; floor(x + 0.5) evaluated in f32 would round odd x to the next even number."""
    n = 16000 * 700
    pcm = np.zeros(n, np.int16)
    s_cs, e_cs = 60000.00625, 60500.0  # 600.0000625 s -> f32 600.00006 -> x16000 = 9600001 (odd, > 2^23)
    mask, out = H.vad_mask_and_merge([(s_cs, e_cs)], pcm)
    x = np.float32(np.float32(mask[0][0]) * np.float32(16000.0))
    assert float(x) == 9600001.0
    assert len(out) == 1 and len(out[0]["samples"]) == int(np.float32(np.float32(mask[0][1]) * np.float32(16000.0))) - 9600001
