"""CPU oracle (test infrastructure only) for wdr_tokenize: whisper.cpp `tokenize()` [UPSTREAM-RECALL], reached through
`initial_prompt` (reference src/transcribe.rs:74-76, 383-386): GPT-2 style word split, then every word is covered from the left by
the longest vocabulary entry matching at the current position; uncovered bytes are skipped; of duplicate strings the highest id wins.
The character classes are ASCII ("C" locale), and matching runs on bytes so that multi-byte UTF-8 falls into the punctuation class."""
import re

_WORD = re.compile(rb"'s|'t|'re|'ve|'m|'ll|'d| ?[A-Za-z]+| ?[0-9]+| ?[^\sA-Za-z0-9]+|\s+(?!\S)|\s+")


def tokenize(token_strings, text):
    """token_strings: list (by id) of bytes / str / None; text: str or bytes -> list of ids."""
    t2i = {}
    for i, t in enumerate(token_strings):
        if t is not None:
            t2i[t if isinstance(t, bytes) else t.encode()] = i
    data = text if isinstance(text, bytes) else text.encode()
    out = []
    for m in _WORD.finditer(data):
        word = m.group(0)
        i, n = 0, len(word)
        while i < n:
            j = n
            while j > i and word[i:j] not in t2i:
                j -= 1
            if j > i:
                out.append(t2i[word[i:j]])
                i = j
            else:
                i += 1
    return out
