"""Host-side mirror of the reference's subtitle formatter `process_segments` (reference src/formatting.rs) — SURVEY §8f row 1: the
CONSUMER of the library's segments / word timestamps.  In production this stays in the crate's Rust; it is restated here so that
library output can be pushed through the same cue logic in tests (leading-space, punctuation and control-token regressions that a
token-id comparison misses) and checked against the reference's own `segments.json`.

Function-for-function with the Rust (same names; file:line in each docstring).  Segments are dicts(start, end, text, words,
speaker_id); words are dicts(text, start, end, probability)."""
import math

try:
    import regex as _regex

    def _graphemes(s):
        return len(_regex.findall(r"\X", s))
except Exception:  # pragma: no cover - regex ships with the image
    def _graphemes(s):
        return len(s)


def round3(x):
    """src/formatting.rs:32-33 (f64::round = half away from zero)."""
    y = x * 1000.0
    return (math.floor(y + 0.5) if y >= 0 else -math.floor(-y + 0.5)) / 1000.0


def default_config():
    """PostProcessConfig::default(), src/formatting.rs:102-120."""
    return dict(max_chars_per_line=38, max_lines=1, cps_cap=17.0, split_gap_sec=0.5, comma_min_chars_before_allow=55, min_word_dur=0.10,
                min_sub_dur=1.0, max_sub_dur=6.0, soft_max_words_per_line=0, insert_interword_space=True, use_grapheme_len=True,
                enforce_kinsoku=False, allow_comma_split=True)


_PROFILES = {  # apply_profile, src/formatting.rs:147-189
    "Latin": dict(max_chars_per_line=38, cps_cap=17.0, insert_interword_space=True, use_grapheme_len=True, enforce_kinsoku=False, allow_comma_split=True),
    "CJK": dict(max_chars_per_line=20, cps_cap=11.5, insert_interword_space=False, use_grapheme_len=True, enforce_kinsoku=True, allow_comma_split=True),
    "SEAsianNoSpace": dict(max_chars_per_line=22, cps_cap=13.0, insert_interword_space=True, use_grapheme_len=True, enforce_kinsoku=False, allow_comma_split=False),
    "RTL": dict(max_chars_per_line=28, cps_cap=14.0, insert_interword_space=True, use_grapheme_len=True, enforce_kinsoku=False, allow_comma_split=True),
    "Indic": dict(max_chars_per_line=30, cps_cap=15.0, insert_interword_space=True, use_grapheme_len=True, enforce_kinsoku=False, allow_comma_split=True),
}


def profile_for_lang(lang):
    """src/formatting.rs:191-204."""
    if lang in ("zh", "zh-CN", "zh-TW", "ja", "ko"):
        return "CJK"
    if lang in ("th", "lo", "km", "my"):
        return "SEAsianNoSpace"
    if lang in ("ar", "fa", "ur", "he"):
        return "RTL"
    if lang in ("hi", "bn", "ta", "te", "ml", "mr", "gu", "pa", "kn", "or", "si"):
        return "Indic"
    return "Latin"


def config_for_language(lang, overrides=None):
    """PostProcessConfig::for_language + apply_overrides (src/formatting.rs:131-133, 54-68; call site src/engine.rs:190-191)."""
    cfg = default_config()
    cfg.update(_PROFILES[profile_for_lang(lang)])
    for k, v in (overrides or {}).items():
        if v is not None:
            assert k in cfg, k
            cfg[k] = v
    return cfg


class VadMaskOracle:
    """src/formatting.rs:217-238: speech intervals; is_silence([t0, t1]) = no overlap with any of them."""

    def __init__(self, mask):
        self.mask = sorted([(s, e) for s, e in mask if e > s], key=lambda t: t[0])

    def is_silence(self, t0, t1):
        if t1 <= t0:
            return True
        for s0, s1 in self.mask:
            if s1 <= t0:
                continue
            if s0 >= t1:
                break
            if s1 > t0 and s0 < t1:
                return False
        return True


class NoSilence:
    def is_silence(self, t0, t1):
        return False


_PUNC_BYTES = set(b'.!?,;:)]}"')  # the byte-wise test of split_trailing_punct only ever matches the ASCII members of its list


def split_trailing_punct(s):
    """src/formatting.rs:358-372 (operates on UTF-8 bytes; `b as char` never equals a multi-byte punctuation mark)."""
    b = s.encode("utf-8")
    cut = len(b)
    for idx in range(len(b) - 1, -1, -1):
        if b[idx] in _PUNC_BYTES:
            cut = idx
        else:
            break
    return (b[:cut].decode("utf-8"), b[cut:].decode("utf-8")) if cut < len(b) else (s, "")


def is_terminal_punct(p):
    return p in (".", "!", "?", "…", "。", "！", "？")


def is_comma_like(p):
    return p in (",", "，", "、", ";")


def is_ascii_word(s):
    return bool(s) and all((c.isascii() and c.isalpha()) or c == "'" for c in s)


def join_tokens(a, b, insert_space):
    """src/formatting.rs:446-456 -> (word, punc, leading_space)."""
    s = a["word"] + a["punc"]
    if insert_space and b["leading_space"] and b["word"] and not s.endswith(" "):
        s += " "
    return s + b["word"], b["punc"], a["leading_space"]


def merge_continuations(toks):
    """src/formatting.rs:321-356."""
    out = []
    for t in toks:
        if out:
            prev = out[-1]
            if not t["word"] and t["punc"]:
                prev["word"], prev["punc"], _ = join_tokens(prev, t, False)
                prev["end"] = max(prev["end"], t["end"])
                continue
            if (not t["leading_space"]) and is_ascii_word(prev["word"]) and is_ascii_word(t["word"]) and not prev["punc"] and (t["start"] - prev["end"]) <= 0.03:
                prev["word"], prev["punc"], _ = join_tokens(prev, t, False)
                prev["end"] = max(prev["end"], t["end"])
                continue
        out.append(t)
    return out


def clamp_and_merge_tiny_words(toks, cfg, oracle):
    """src/formatting.rs:380-444."""
    if not toks:
        return toks
    n = len(toks)
    for i in range(n):
        dur = toks[i]["end"] - toks[i]["start"]
        if dur < cfg["min_word_dur"]:
            grow = (cfg["min_word_dur"] - dur) / 2.0
            toks[i]["start"] -= grow
            toks[i]["end"] += grow
        if i > 0:
            mid = 0.5 * (toks[i - 1]["end"] + toks[i]["start"])
            toks[i - 1]["end"] = min(toks[i - 1]["end"], mid)
            toks[i]["start"] = max(toks[i]["start"], mid)
        if i + 1 < n:
            mid = 0.5 * (toks[i]["end"] + toks[i + 1]["start"])
            toks[i]["end"] = min(toks[i]["end"], mid)
            toks[i + 1]["start"] = max(toks[i + 1]["start"], mid)
        pad = 0.02
        if oracle.is_silence(toks[i]["start"] - pad, toks[i]["start"]):
            toks[i]["start"] += pad
        if oracle.is_silence(toks[i]["end"], toks[i]["end"] + pad):
            toks[i]["end"] -= pad
    out, i = [], 0
    while i < n:
        dur = toks[i]["end"] - toks[i]["start"]
        if dur < cfg["min_word_dur"] and i + 1 < n:
            nxt = dict(toks[i + 1])
            w, p, ls = join_tokens(toks[i], nxt, cfg["insert_interword_space"])
            nxt.update(word=w, punc=p, start=min(toks[i]["start"], nxt["start"]), leading_space=ls)
            out.append(nxt)
            i += 2
        elif dur < cfg["min_word_dur"] and i > 0:
            prev = out.pop()
            w, p, ls = join_tokens(prev, toks[i], cfg["insert_interword_space"])
            prev.update(word=w, punc=p, end=max(prev["end"], toks[i]["end"]), leading_space=ls)
            out.append(prev)
            i += 1
        else:
            out.append(dict(toks[i]))
            i += 1
    return out


def split_into_groups(toks, cfg):
    """src/formatting.rs:458-472."""
    groups, cur = [], []
    for i, t in enumerate(toks):
        cur.append(t)
        long_gap = i + 1 < len(toks) and (toks[i + 1]["start"] - t["end"]) >= cfg["split_gap_sec"]
        if is_terminal_punct(t["punc"]) or long_gap:
            groups.append(cur)
            cur = []
    if cur:
        groups.append(cur)
    return groups


def slice_chars(sl, cfg):
    """src/formatting.rs:614-622."""
    if cfg["use_grapheme_len"]:
        core = sum(_graphemes(t["word"]) + _graphemes(t["punc"]) for t in sl)
    else:
        core = sum(len(t["word"].encode("utf-8")) + len(t["punc"].encode("utf-8")) for t in sl)
    spaces = sum(1 for t in sl[1:] if t["leading_space"]) if cfg["insert_interword_space"] else 0
    return core + spaces


def render_slice(sl, cfg):
    """src/formatting.rs:604-612."""
    s = ""
    for i, t in enumerate(sl):
        if cfg["insert_interword_space"] and t["leading_space"] and i > 0:
            s += " "
        s += t["word"] + t["punc"]
    return s


def length_penalty(chars, cap):
    return 0.0 if chars <= cap else 0.02 * float(chars - cap) ** 2


def soft_cap_penalty(v, cap):
    return 0.0 if v <= cap else 0.01 * float(v - cap) ** 2


_SHORT_FUNCT = ("i", "to", "a", "the", "and", "or", "of", "in", "on", "for", "with", "at")


def syntax_penalty(left, right):
    """src/formatting.rs:632-648."""
    rw, lw = right.split(), left.split()
    pen = 0.0
    if rw and rw[0].lower() in _SHORT_FUNCT:
        pen += 0.3
    if lw and lw[-1].lower() in _SHORT_FUNCT:
        pen += 0.25
    return pen


def split_into_lines(sl, cfg):
    """src/formatting.rs:521-602."""
    if not sl:
        return [""]
    if cfg["max_lines"] <= 1:
        return [render_slice(sl, cfg)]
    if slice_chars(sl, cfg) <= cfg["max_chars_per_line"]:
        return [render_slice(sl, cfg)]
    cands = []
    for k in range(1, len(sl)):
        left_term = sl[k - 1]["punc"]
        long_gap = (sl[k]["start"] - sl[k - 1]["end"]) >= cfg["split_gap_sec"]
        comma_ok = is_comma_like(left_term) and slice_chars(sl, cfg) >= cfg["comma_min_chars_before_allow"]
        if is_terminal_punct(left_term) or long_gap or comma_ok or k % 2 == 0 or k == len(sl) // 2:
            cands.append(k)
    if not cands:
        return [render_slice(sl, cfg)]
    best_k, best_score = cands[0], math.inf
    for k in cands:
        lchars, rchars = slice_chars(sl[:k], cfg), slice_chars(sl[k:], cfg)
        ltext, rtext = render_slice(sl[:k], cfg), render_slice(sl[k:], cfg)
        len_pen = length_penalty(lchars, cfg["max_chars_per_line"]) + length_penalty(rchars, cfg["max_chars_per_line"])
        word_pen = 0.0
        if cfg["soft_max_words_per_line"] > 0:
            word_pen = soft_cap_penalty(k, cfg["soft_max_words_per_line"]) + soft_cap_penalty(len(sl) - k, cfg["soft_max_words_per_line"])
        left_term = sl[k - 1]["punc"]
        gap = sl[k]["start"] - sl[k - 1]["end"]
        bonus = -0.6 * float(is_terminal_punct(left_term)) + -0.3 * float(gap >= cfg["split_gap_sec"]) + 0.15 * float(is_comma_like(left_term))
        cont_pen = 5.0 if not sl[k]["leading_space"] else 0.0
        score = len_pen + word_pen + syntax_penalty(ltext, rtext) + bonus + cont_pen
        if score < best_score:
            best_score, best_k = score, k
    return [render_slice(sl[:best_k], cfg), render_slice(sl[best_k:], cfg)]


def slice_stats(sl, cfg):
    t0 = sl[0]["start"] if sl else 0.0
    t1 = sl[-1]["end"] if sl else t0
    return t0, t1, slice_chars(sl, cfg)


def build_cue(group, start_idx, cfg):
    """src/formatting.rs:474-512 -> (next index, cue segment)."""
    j = start_idx + 1
    while True:
        t0, t1, chars = slice_stats(group[start_idx:j], cfg)
        dur = max(t1 - t0, 0.001)
        cps = chars / dur
        if j < len(group) and dur < cfg["max_sub_dur"] and (cps <= cfg["cps_cap"] or chars < cfg["max_chars_per_line"] * cfg["max_lines"]):
            j += 1
        else:
            break
    sl = group[start_idx:j]
    t0, t1, _ = slice_stats(sl, cfg)
    words = [dict(text=t["word"] + t["punc"], start=round3(t["start"]), end=round3(t["end"]), probability=t["prob"]) for t in sl]
    cue = dict(start=round3(max(t0, 0.0)), end=round3(t1), text="\n".join(split_into_lines(sl, cfg)), words=words, speaker_id=sl[0]["speaker"])
    return j, cue


def process_segments(segments, cfg, oracle=None):
    """src/formatting.rs:240-313: whisper segments (with word/token spans) -> subtitle cues."""
    oracle = oracle or NoSilence()
    allw = []
    for seg in segments:
        if seg.get("words") is not None:
            allw += [(seg.get("speaker_id"), w) for w in seg["words"]]
        elif seg["text"].strip():
            allw.append((seg.get("speaker_id"), dict(text=seg["text"], start=seg["start"], end=seg["end"], probability=None)))
    if not allw:
        return []
    toks = []
    for speaker, w in allw:
        core_raw, punc = split_trailing_punct(w["text"])
        leading_space = core_raw.startswith(" ") or core_raw.startswith("\n")
        core = core_raw.lstrip(" \n").replace("�", "")
        punc = punc.replace("�", "")
        if not core and not punc:
            continue
        toks.append(dict(word=core, punc=punc, start=float(w["start"]), end=float(w["end"]), prob=w.get("probability"), speaker=speaker,
                         leading_space=leading_space))
    toks = merge_continuations(toks)
    toks = clamp_and_merge_tiny_words(toks, cfg, oracle)
    cues = []
    for g in split_into_groups(toks, cfg):
        i = 0
        while i < len(g):
            i, cue = build_cue(g, i, cfg)
            cues.append(cue)
    return cues
