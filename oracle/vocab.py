"""Synthetic Whisper vocabulary (test infrastructure): special-token ids as whisper.cpp derives them from n_vocab
 and deterministic token strings.  No tokenizer file can exist offline, so text ids map to short
pseudo-words that exercise leading spaces, punctuation and digits (voice_length weights); specials use
whisper.cpp's bracket names, which the reference's control-token filter relies on (src/transcribe.rs:206-212).
csrc/vocab.cuh restates the same mapping for the library."""

LANGS = ["en", "zh", "de", "es", "ru", "ko", "fr", "ja", "pt", "tr", "pl", "ca", "nl", "ar", "sv", "it", "id", "hi", "fi", "vi", "he", "uk",
         "el", "ms", "cs", "ro", "da", "hu", "ta", "no", "th", "ur", "hr", "bg", "lt", "la", "mi", "ml", "cy", "sk", "te", "fa", "lv", "bn",
         "sr", "az", "sl", "kn", "et", "mk", "br", "eu", "is", "hy", "ne", "mn", "bs", "kk", "sq", "sw", "gl", "mr", "pa", "si", "km", "sn",
         "yo", "so", "af", "oc", "ka", "be", "tg", "sd", "gu", "am", "yi", "lo", "uz", "fo", "ht", "ps", "tk", "nn", "mt", "sa", "lb", "my",
         "bo", "tl", "mg", "as", "tt", "haw", "ln", "ha", "ba", "jw", "su", "yue"]


def special_ids(n_vocab):
    ml = n_vocab >= 51865
    eot = 50256 + (1 if ml else 0)
    sot = eot + 1
    beg = n_vocab - 1501
    v = dict(n_vocab=n_vocab, eot=eot, sot=sot, translate=beg - 6, transcribe=beg - 5, solm=beg - 4, prev=beg - 3, nosp=beg - 2,
             not_=beg - 1, beg=beg, lang0=sot + 1, n_langs=100, space=220, multilingual=ml)
    return v


def voice_length(text):
    import numpy as np
    r = np.float32(0.0)
    for c in text:
        if c == " ":
            r = np.float32(r + np.float32(0.01))
        elif c == ",":
            r = np.float32(r + np.float32(2.0))
        elif c in ".!?":
            r = np.float32(r + np.float32(3.0))
        elif "0" <= c <= "9":
            r = np.float32(r + np.float32(3.0))
        else:
            r = np.float32(r + np.float32(1.0))
    return float(r)


def token_text(i, n_vocab):
    v = special_ids(n_vocab)
    if i < v["eot"]:
        if i == v["space"]:
            return " "
        s, n = "", i
        while True:
            s += chr(ord("a") + n % 26)
            n //= 26
            if n == 0:
                break
        pre = " " if i % 3 == 0 else ""
        suf = "," if i % 17 == 0 else "." if i % 29 == 0 else "?" if i % 31 == 0 else str(i % 10) if i % 23 == 0 else ""
        return pre + s + suf
    if i == v["eot"]:
        return "[_EOT_]"
    if i == v["sot"]:
        return "[_SOT_]"
    if i == v["translate"]:
        return "[_TRANSLATE_]"
    if i == v["transcribe"]:
        return "[_TRANSCRIBE_]"
    if i == v["solm"]:
        return "[_SOLM_]"
    if i == v["prev"]:
        return "[_PREV_]"
    if i == v["nosp"]:
        return "[_NOSP_]"
    if i == v["not_"]:
        return "[_NOT_]"
    if i == v["beg"]:
        return "[_BEG_]"
    if i > v["beg"]:
        return f"[_TT_{i - v['beg']}]"
    if v["lang0"] <= i < v["lang0"] + (n_vocab - 51765 - (1 if v["multilingual"] else 0)):
        return f"[_LANG_{LANGS[i - v['lang0']]}]"
    return f"[_extra_token_{i}]"
