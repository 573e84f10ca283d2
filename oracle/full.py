"""whisper_full_with_state as the reference drives it, composed from the oracle's C pieces, for ONE
buffer of <= 30 s (the sharded mode of SURVEY §0.4: single_segment, no prompt carry, per-chunk mel max).  Test
infrastructure only."""
import numpy as np

from . import native, vocab as V, weights as W, filters


def prompt_tokens(n_vocab, lang_id=0, translate=False):
    v = V.special_ids(n_vocab)
    p = [v["sot"]]
    if v["multilingual"]:
        p += [v["lang0"] + lang_id, v["translate"] if translate else v["transcribe"]]
    return p


def dtw_sequence(n_vocab, text_ids, lang_id=0):
    v = V.special_ids(n_vocab)
    seq = [v["sot"]]
    if v["multilingual"]:
        seq.append(v["lang0"] + lang_id)
    sot_len = len(seq)
    seq.append(v["not_"])
    seq += [int(t) for t in text_ids]
    seq.append(v["eot"])
    return seq, sot_len


def n_len_org(n_samples):
    return 1 + (n_samples + 200 - 400) // 160


def detect_language(dec, enc_out):
    """whisper_lang_auto_detect: decode [SOT] at position 0 against the window, arg-max over the language tokens."""
    nv = dec.a["n_vocab"]
    v = V.special_ids(nv)
    dec.set_audio(enc_out)
    lg = dec.step(v["sot"], 0)
    n_l = v["translate"] - v["lang0"]
    return int(np.argmax(lg[v["lang0"]: v["lang0"] + n_l]))


def full_window(dec, enc_out, pcm_f32, dtw=True, token_timestamps=True, delta_min=10, lang_id=0):
    """dec: native.Decoder; enc_out [1500, d] (encoder output for this window); pcm_f32: the window's samples (<= 480000).
    Returns dict(segments=[dict(t0, t1, text, tokens=[TokenData])], seek_delta, no_speech_prob, margins, ...)."""
    nv = dec.a["n_vocab"]
    v = V.special_ids(nv)
    seek, seek_end = 0, n_len_org(len(pcm_f32))
    out = dict(segments=[], seek_end=seek_end)
    if seek_end < seek + delta_min or seek + delta_min >= seek_end:
        return out
    dec.set_audio(enc_out)
    r = dec.decode_window(prompt_tokens(nv, lang_id), seek, seek_end, True, delta_min)
    out.update(r)
    toks = r["tokens"]
    if r["failed"]:
        avg_logprob = -np.inf
    else:
        avg_logprob = sum(float(t.plog) for t in toks) / max(1, r["result_len"]) if r["result_len"] else -np.inf
    is_no_speech = r["no_speech_prob"] > 0.6 and avg_logprob < -1.0
    out["is_no_speech"] = bool(is_no_speech)
    if not toks or is_no_speech:
        return out
    seek_delta = r["seek_delta"]
    t0 = seek + 2 * (toks[0].tid - v["beg"])
    text = "".join(V.token_text(t.id, nv) for t in toks if t.id < v["eot"])
    if text:
        t1 = seek + seek_delta
        if token_timestamps:
            energy = native.signal_energy(pcm_f32, 32)
            vlen = [V.voice_length(V.token_text(t.id, nv)) for t in toks]
            st3 = np.zeros(3, np.int64)
            toks = native.token_timestamps(toks, t0, t1, vlen, energy, v["beg"], v["eot"], st3)
        if dtw:
            n_frames = min(3000, seek_delta, seek_end - seek)
            n_audio = n_frames // 2
            text_ids = [t.id for t in toks if t.id < v["eot"]]
            seq, sot_len = dtw_sequence(nv, text_ids, lang_id)
            aheads = W.ALIGNMENT_HEADS[dec.arch]
            w = dec.dtw_attention(seq, aheads, n_audio)
            cost = native.dtw_cost(w, sot_len, 7)
            ti, tj = native.dtw(cost)
            toks = native.dtw_stamp(toks, v["eot"], ti, tj, seek)
            out.update(dtw_w=w, dtw_cost=cost, dtw_path=(ti, tj))
        out["segments"].append(dict(t0=t0, t1=t1, text=text, tokens=toks))
    return out


def full_sequential(dec, encode, filt, pcm_f32, dtw=True, token_timestamps=True, delta_min=10, n_max_text_ctx=16384, prompt_past=None):
    """whisper_full_with_state's seek loop over a buffer of any length: the path the crate takes with VAD and
    diarization off (one SpeechSegment = the whole file, reference src/engine.rs:124-134 -> state.full, src/transcribe.rs:389).
    greedy T=0, single_segment, no_context (prompt_past starts empty and accumulates across the call's windows), global mel max.

    dec: native.Decoder; encode(mel_window [n_mel, 3000], seek) -> [1500, d] (seek lets a test substitute the encoder output the
    library computed for the same window); filt: mel filterbank.
    Returns dict(segments=[dict(t0, t1, text, tokens)], windows=[dict(seek, seek_delta, n_sampled, failed, result_len)])."""
    nv = dec.a["n_vocab"]
    v = V.special_ids(nv)
    n = len(pcm_f32)
    mel = native.log_mel(pcm_f32, filt)  # [n_mel, (n + 480000) // 160], normalised with the buffer-global max
    seek, seek_end = 0, n_len_org(n)
    out = dict(segments=[], windows=[])
    if seek_end < seek + delta_min:
        return out
    energy = native.signal_energy(pcm_f32, 32) if token_timestamps else None
    st3 = np.zeros(3, np.int64)
    past = list(prompt_past or [])
    while seek + delta_min < seek_end:
        win = np.zeros((mel.shape[0], 3000), np.float32)
        take = max(0, min(3000, mel.shape[1] - seek))
        win[:, :take] = mel[:, seek:seek + take]
        dec.set_audio(encode(win, seek))
        prompt = []
        n_take = 0
        if past and n_max_text_ctx > 0:
            n_take = min(n_max_text_ctx, 448 // 2, len(past))
            prompt = [v["prev"]] + past[len(past) - n_take:]
        prompt += prompt_tokens(nv)
        r = dec.decode_window(prompt, seek, seek_end, True, delta_min)
        toks = r["tokens"]
        out["windows"].append(dict(seek=seek, seek_delta=r["seek_delta"], n_sampled=r["n_sampled"], failed=r["failed"], result_len=r["result_len"]))
        if r["failed"]:
            avg_logprob = -np.inf
        else:
            avg_logprob = sum(float(t.plog) for t in toks) / max(1, r["result_len"]) if r["result_len"] else -np.inf
        is_no_speech = r["no_speech_prob"] > 0.6 and avg_logprob < -1.0
        new_ids = []
        if toks and not is_no_speech:
            seek_delta = r["seek_delta"]
            t0 = seek + 2 * (toks[0].tid - v["beg"])
            text = "".join(V.token_text(t.id, nv) for t in toks if t.id < v["eot"])
            if text:
                t1 = seek + seek_delta
                if token_timestamps:
                    vlen = [V.voice_length(V.token_text(t.id, nv)) for t in toks]
                    toks = native.token_timestamps(toks, t0, t1, vlen, energy, v["beg"], v["eot"], st3)
                if dtw:
                    n_frames = min(3000, seek_delta, seek_end - seek)
                    text_ids = [t.id for t in toks if t.id < v["eot"]]
                    seq, sot_len = dtw_sequence(nv, text_ids)
                    w = dec.dtw_attention(seq, W.ALIGNMENT_HEADS[dec.arch], n_frames // 2)
                    ti, tj = native.dtw(native.dtw_cost(w, sot_len, 7))
                    toks = native.dtw_stamp(toks, v["eot"], ti, tj, seek)
                out["segments"].append(dict(t0=t0, t1=t1, text=text, tokens=toks))
                new_ids = [int(t.id) for t in toks[: r["result_len"]]]
        past = (past[len(past) - n_take:] if n_take else []) + new_ids
        if r["seek_delta"] <= 0:
            break
        seek += r["seek_delta"]
    return out
