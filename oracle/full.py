"""whisper_full_with_state as the reference drives it, composed from the oracle's C pieces, for ONE
buffer of <= 30 s (the sharded mode of SURVEY §0.4: single_segment, no prompt carry, per-chunk mel max).  Test
infrastructure only."""
import math

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


def sequence_entropy(tokens, result_len):
    """whisper_sequence_score: entropy of the ids of the last 32 kept tokens (std::map order = ascending id)."""
    ids = [t.id for t in tokens[max(0, result_len - 32): result_len]]
    e = 0.0
    for i in sorted(set(ids)):
        p = ids.count(i) / len(ids)
        e -= p * math.log(p)
    return e


def beam_decode_window(dec, prompt, seek, seek_end, beam_size=5, delta_min=10, suppress_blank=True, max_initial_ts=1.0,
                       n_decoders=None, temperature=0.0, entropy_thold=2.4):
    """whisper_full's beam search for one window at one temperature (the crate's default strategy, reference src/transcribe.rs:22,
    29-32), restated from whisper.cpp: logits are divided by the temperature when it is > 0 (whisper_process_logits); every live
    decoder proposes its `beam_size` best tokens (by log-probability, ties to the lower id); candidates are stably sorted by
    cumulative log-probability; each live decoder, in index order, takes the next candidate and — after the first iteration — skips
    the candidates that repeat its token sequence; the per-decoder bookkeeping is the greedy loop's (timestamp pairing, seek_delta,
    completion, failure); ranking (whisper_sequence_score): failed decoders are skipped, a decoder whose last 32 kept tokens have
    entropy < entropy_thold (result_len > 32) fails, the best average log-probability wins (first maximum, decoder 0 if none).
    n_decoders: beam_size at temperature 0, max(1, greedy.best_of) = 1 above (whisper_full_default_params leaves best_of at -1 for
    the beam strategy).  dec.set_audio(...) must have been called.  Returns the dict decode_window returns (tokens = the winner's)
    plus dec_failed (the winner counts as failed for the fallback test) and avg_logprob."""
    nv = dec.a["n_vocab"]
    v = V.special_ids(nv)
    n_prompt = len(prompt)
    n_max = 448 // 2 - 4
    n_dec = beam_size if n_decoders is None else n_decoders
    T = np.float32(temperature)

    def scaled(lg):
        return (lg / T).astype(np.float32) if temperature > 0 else lg

    for i, t in enumerate(prompt):
        lg0 = dec.step(int(t), i, want_logits=(i == n_prompt - 1))
    kv0 = dec.get_kv()
    ls0 = lg0  # whisper_full reads no_speech_prob from the RAW logits of the prompt decode, before the temperature division
    lp0 = ls0 - (np.log(np.exp((ls0 - ls0.max()).astype(np.float32)).sum(dtype=np.float32)) + ls0.max())
    no_speech_prob = float(np.exp(np.float32(lp0[v["nosp"]])))

    class Beam:
        pass

    beams = []
    for k in range(n_dec):
        b = Beam()
        b.tokens, b.sum_all, b.has_ts, b.seek_delta, b.result_len, b.failed, b.completed = [], 0.0, 0, 3000, 0, False, False
        b.logits, b.kv = lg0, kv0
        beams.append(b)
    for i in range(n_max):
        cl = []
        props = {}
        for k, b in enumerate(beams):
            if b.completed or b.failed:
                continue
            _, _, lpb, pb = native.process_logits(scaled(b.logits), nv, [t.id for t in b.tokens], b.has_ts, b.seek_delta, suppress_blank, max_initial_ts)
            fin = np.flatnonzero(lpb > -np.inf)
            order = fin[np.lexsort((fin, -lpb[fin].astype(np.float64)))][:beam_size]
            g = native.sample_stats(pb, lpb, nv)  # tid / pt / ptsum of the distribution
            props[k] = (order, lpb, pb, g)
            for r, tid_ in enumerate(order):
                cl.append((k, r, b.sum_all + float(lpb[tid_])))
        if not cl:
            break
        cl.sort(key=lambda c: -c[2])  # Python's sort is stable: generation order (beam, rank) breaks ties, as std::stable_sort does

        def tok_of(c):
            return int(props[c[0]][0][c[1]])

        def same_seq(a, b_):
            if tok_of(a) != tok_of(b_):
                return False
            ta, tb = beams[a[0]].tokens, beams[b_[0]].tokens
            return len(ta) == len(tb) and all(x.id == y.id for x, y in zip(ta, tb))

        new = list(beams)
        cur_c = 0
        any_live = False
        for k, b in enumerate(beams):
            if b.completed or b.failed:
                continue
            if cur_c >= len(cl):
                cur_c = 0
            cur = cl[cur_c]
            cur_c += 1
            while len(cl) > cur_c and i > 0 and same_seq(cl[cur_c], cur):
                cur_c += 1
            src = beams[cur[0]]
            order, lpb, pb, g = props[cur[0]]
            tid_ = int(order[cur[1]])
            td = native.TokenData()
            td.id, td.tid, td.p, td.plog, td.pt, td.ptsum, td.t0, td.t1, td.t_dtw, td.vlen = tid_, g.tid, float(pb[tid_]), float(lpb[tid_]), g.pt, g.ptsum, -1, -1, -1, 0.0
            if tid_ >= v["beg"]:
                td.tid, td.pt = tid_, td.p
            h = Beam()
            h.tokens, h.sum_all, h.has_ts, h.seek_delta, h.result_len = src.tokens + [td], cur[2], src.has_ts, src.seek_delta, src.result_len
            h.failed, h.completed, h.logits, h.kv = False, False, None, src.kv
            done = False
            if td.id > v["beg"]:
                sd_new = 2 * (td.id - v["beg"])
                if h.has_ts and h.seek_delta > sd_new and h.result_len < i:
                    h.failed, done = True, True
                else:
                    h.seek_delta, h.result_len, h.has_ts = sd_new, i + 1, 1
            if not done and (td.id == v["eot"] or (h.has_ts and seek + h.seek_delta + delta_min >= seek_end)):
                fail = False
                if h.result_len == 0:
                    if seek + h.seek_delta + delta_min >= seek_end:
                        h.result_len = i + 1
                    else:
                        fail = True
                if fail:
                    h.failed = True
                else:
                    h.result_len, h.seek_delta, h.completed = i + 1, 3000, True  # single_segment
                done = True
            if not done and i == n_max - 1 and (h.result_len == 0 or h.seek_delta < 3000 // 2):
                h.failed, done = True, True
            new[k] = h
            any_live = any_live or not done
        beams = new
        if not any_live or i == n_max - 1:
            break
        for b in beams:
            if b.completed or b.failed:
                continue
            dec.set_kv(b.kv)
            b.logits = dec.step(b.tokens[-1].id, n_prompt + i)
            b.kv = dec.get_kv()
    best, best_score, failed_k = 0, -np.inf, [b.failed for b in beams]
    for k, b in enumerate(beams):
        if b.failed:
            continue
        if b.result_len <= 0:
            failed_k[k] = True
            continue
        score = sum(float(t.plog) for t in b.tokens[: b.result_len]) / b.result_len
        if b.result_len > 32 and sequence_entropy(b.tokens, b.result_len) < entropy_thold:
            failed_k[k] = True
            continue
        if best_score < score:
            best, best_score = k, score
    w = beams[best]
    n_keep = len(w.tokens) if w.failed else w.result_len
    avg = sum(float(t.plog) for t in w.tokens[: w.result_len]) / w.result_len if w.result_len > 0 else -np.inf
    return dict(tokens=w.tokens[:n_keep], seek_delta=w.seek_delta, failed=w.failed, completed=w.completed, n_sampled=len(w.tokens),
                result_len=w.result_len, no_speech_prob=no_speech_prob, margins=np.zeros(0, np.float32), beam=best,
                dec_failed=bool(failed_k[best]), avg_logprob=avg)


def temperature_ladder(temperature_inc, temperature=0.0):
    """whisper_full: for (float t = temperature; t < 1.0f + 1e-6f; t += temperature_inc) — float32 arithmetic."""
    if temperature_inc <= 0:
        return [float(temperature)]
    out, t, inc = [], np.float32(temperature), np.float32(temperature_inc)
    while t < np.float32(1.0) + np.float32(1e-6):
        out.append(float(t))
        t = np.float32(t + inc)
    return out


def full_window(dec, enc_out, pcm_f32, dtw=True, token_timestamps=True, delta_min=10, lang_id=0, beam_size=1,
                temperature_inc=0.0, logprob_thold=-1.0, no_speech_thold=0.6, entropy_thold=2.4, temperature=0.0, translate=False):
    """dec: native.Decoder; enc_out [1500, d] (encoder output for this window); pcm_f32: the window's samples (<= 480000).
    Returns dict(segments=[dict(t0, t1, text, tokens=[TokenData])], seek_delta, no_speech_prob, margins, ...)."""
    nv = dec.a["n_vocab"]
    v = V.special_ids(nv)
    seek, seek_end = 0, n_len_org(len(pcm_f32))
    out = dict(segments=[], seek_end=seek_end)
    if seek_end < seek + delta_min or seek + delta_min >= seek_end:
        return out
    dec.set_audio(enc_out)
    temps = temperature_ladder(temperature_inc, temperature)
    assert (len(temps) == 1 and temps[0] <= 0) or beam_size > 1, "above T = 0 only the beam strategy is restated (deterministic: one decoder)"
    for it, t_cur in enumerate(temps):
        if beam_size > 1:
            r = beam_decode_window(dec, prompt_tokens(nv, lang_id, translate), seek, seek_end, beam_size, delta_min,
                                   n_decoders=beam_size if t_cur <= 0 else 1, temperature=t_cur, entropy_thold=entropy_thold)
        else:
            r = dec.decode_window(prompt_tokens(nv, lang_id, translate), seek, seek_end, True, delta_min)
        r["temperature"] = t_cur
        # "was the decoding successful for the current temperature?" — the last temperature's result stands whatever it is
        if it != len(temps) - 1 and (r["dec_failed"] or r["result_len"] <= 0 or
                                     (r["avg_logprob"] < logprob_thold and r["no_speech_prob"] < no_speech_thold)):
            continue
        break
    out.update(r)
    toks = r["tokens"]
    if r["failed"]:
        avg_logprob = -np.inf
    else:
        avg_logprob = sum(float(t.plog) for t in toks) / max(1, r["result_len"]) if r["result_len"] else -np.inf
    is_no_speech = r["no_speech_prob"] > no_speech_thold and avg_logprob < logprob_thold
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
        # whisper_full appends tokens_cur[0 .. result_len) to prompt_past unless the window is a no-speech window, whether or not a
        # segment comes out of them
        new_ids = [] if is_no_speech else [int(t.id) for t in toks[: r["result_len"]]]
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
        past = (past[len(past) - n_take:] if n_take else []) + new_ids
        if r["seek_delta"] <= 0:
            break
        seek += r["seek_delta"]
    return out
