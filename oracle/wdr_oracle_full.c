/*
 * wdr_oracle_full.c — CPU restatement of the decoder half of whisper.cpp's whisper_full_with_state as the
 * reference drives it (reference src/transcribe.rs:20-87 setup_params: token_timestamps, single_segment,
 * suppress_blank; :389 state.full; :252-282 token accessors): cross-KV projection, KV-cached decoder step,
 * whisper_process_logits, greedy whisper_sample_token, the per-window decode loop, the heuristic token
 * timestamps (whisper_exp_compute_token_level_timestamps) and the DTW token timestamps
 * (whisper_exp_compute_token_level_timestamps_dtw), following SURVEY Appendix A.3-A.6.
 *
 * TEST INFRASTRUCTURE ONLY (see wdr_oracle.c header).  PARITY UNPINNED: whisper.cpp is an un-vendored
 * dependency (whisper-rs 0.15.0 -> whisper.cpp ~v1.7.6); the decoder arithmetic is cross-checked against
 * transformers' WhisperDecoder in tests/test_oracle_decoder.py.
 *
 * Precision policy switch `bf16`: 0 = fp32 everywhere (the whisper.cpp-fp32 gold the north-star tolerances refer to);
 * 1 = the storage precision of libwdr_b200's decoder: the encoder output and the cross-KV cache are rounded to bf16 once
 * and the self-KV cache to f16 when a position is appended (whisper.cpp keeps both KV caches in f16); every activation and all
 * accumulation stay fp32 (the library carries activations as (hi, lo) bf16 pairs = 16 mantissa bits, see csrc/decoder.cu).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

typedef struct {
    int32_t id, tid;
    float p, plog, pt, ptsum;
    int64_t t0, t1, t_dtw;
    float vlen;
} oracle_token_data; /* == whisper_token_data */

typedef struct {
    int n_vocab, eot, sot, translate, transcribe, solm, prev, nosp, not_, beg, lang0, n_langs, space;
} oracle_vocab;

typedef struct {
    int d, n_head, n_layer, n_vocab, bf16;
    const float *tok_emb, *pos;
    const float **lw; /* per layer: 24 pointers */
    const float *ln_g, *ln_b;
    float *ck, *cv;   /* [L][1500][d] */
    float *sk, *sv;   /* [L][448][d] */
    float *scratch;
} oracle_dec;

enum { W_LN1G, W_LN1B, W_QW, W_QB, W_KW, W_VW, W_VB, W_OW, W_OB, W_LN2G, W_LN2B, W_CQW, W_CQB, W_CKW, W_CVW, W_CVB, W_COW, W_COB,
       W_LN3G, W_LN3B, W_F1W, W_F1B, W_F2W, W_F2B, W_COUNT };

static inline float bf16r(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    if ((u & 0x7f800000u) == 0x7f800000u) return x;
    u = (u + 0x7fffu + ((u >> 16) & 1u)) & 0xffff0000u;
    memcpy(&x, &u, 4);
    return x;
}
static inline float rnd(const oracle_dec *m, float x) { return m->bf16 ? bf16r(x) : x; }
/* float -> IEEE binary16 (round-to-nearest-even, overflow to inf, gradual underflow) -> float: what storing into an f16 cache does */
static inline float f16r(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    const uint32_t a = u & 0x7fffffffu;
    if (a >= 0x7f800000u) return x;                          /* inf / nan */
    if (a < 0x38800000u) return rintf(x * 16777216.0f) / 16777216.0f; /* |x| < 2^-14: multiples of 2^-24 (rintf: ties to even) */
    u = (u + 0xfffu + ((u >> 13) & 1u)) & 0xffffe000u;
    if ((u & 0x7fffffffu) >= 0x47800000u) u = (u & 0x80000000u) | 0x7f800000u; /* >= 65520 rounds to inf */
    memcpy(&x, &u, 4);
    return x;
}
static inline float rnd16(const oracle_dec *m, float x) { return m->bf16 ? f16r(x) : x; }
static inline float gelu_tanh_f(float x) {
    return 0.5f * x * (1.0f + tanhf(0.79788456080286535587989211986876f * x * (1.0f + 0.044715f * x * x)));
}

/* y[N] = W[N][K] x[K] + b */
static void gemv(const float *W, const float *b, const float *x, float *y, int N, int K) {
#pragma omp parallel for schedule(static)
    for (int j = 0; j < N; j++) {
        const float *w = W + (size_t)j * K;
        float s = 0.0f;
#pragma omp simd reduction(+ : s)
        for (int k = 0; k < K; k++) s += w[k] * x[k];
        y[j] = s + (b ? b[j] : 0.0f);
    }
}

static void layer_norm(const float *x, const float *g, const float *b, float *y, int d) {
    double s = 0.0;
    for (int i = 0; i < d; i++) s += x[i];
    const float mean = (float)(s / d);
    double q = 0.0;
    for (int i = 0; i < d; i++) { const float v = x[i] - mean; q += (double)v * v; }
    const float rstd = 1.0f / sqrtf((float)(q / d) + 1e-5f);
    for (int i = 0; i < d; i++) y[i] = (x[i] - mean) * rstd * g[i] + b[i];
}

oracle_dec *oracle_dec_create(int d, int n_head, int n_layer, int n_vocab, const float *wts, int bf16) {
    oracle_dec *m = (oracle_dec *)calloc(1, sizeof(oracle_dec));
    m->d = d; m->n_head = n_head; m->n_layer = n_layer; m->n_vocab = n_vocab; m->bf16 = bf16;
    const float *p = wts;
    m->tok_emb = p; p += (size_t)n_vocab * d;
    m->pos = p; p += (size_t)448 * d;
    m->lw = (const float **)malloc(sizeof(float *) * W_COUNT * n_layer);
    const size_t dd = (size_t)d * d;
    for (int l = 0; l < n_layer; l++) {
        const float **w = m->lw + (size_t)l * W_COUNT;
        w[W_LN1G] = p; p += d; w[W_LN1B] = p; p += d;
        w[W_QW] = p; p += dd; w[W_QB] = p; p += d;
        w[W_KW] = p; p += dd;
        w[W_VW] = p; p += dd; w[W_VB] = p; p += d;
        w[W_OW] = p; p += dd; w[W_OB] = p; p += d;
        w[W_LN2G] = p; p += d; w[W_LN2B] = p; p += d;
        w[W_CQW] = p; p += dd; w[W_CQB] = p; p += d;
        w[W_CKW] = p; p += dd;
        w[W_CVW] = p; p += dd; w[W_CVB] = p; p += d;
        w[W_COW] = p; p += dd; w[W_COB] = p; p += d;
        w[W_LN3G] = p; p += d; w[W_LN3B] = p; p += d;
        w[W_F1W] = p; p += 4 * dd; w[W_F1B] = p; p += 4 * d;
        w[W_F2W] = p; p += 4 * dd; w[W_F2B] = p; p += d;
    }
    m->ln_g = p; p += d;
    m->ln_b = p; p += d;
    m->ck = (float *)malloc(sizeof(float) * (size_t)n_layer * 1500 * d);
    m->cv = (float *)malloc(sizeof(float) * (size_t)n_layer * 1500 * d);
    m->sk = (float *)calloc((size_t)n_layer * 448 * d, sizeof(float));
    m->sv = (float *)calloc((size_t)n_layer * 448 * d, sizeof(float));
    m->scratch = (float *)malloc(sizeof(float) * (size_t)(16 * d + 1500 * n_head + 448 * n_head));
    return m;
}

void oracle_dec_free(oracle_dec *m) {
    if (!m) return;
    free((void *)m->lw); free(m->ck); free(m->cv); free(m->sk); free(m->sv); free(m->scratch); free(m);
}

/* self-attention cache snapshots (beam search: a beam that switches parents continues from the parent's cache) */
void oracle_dec_get_self_kv(const oracle_dec *m, float *k, float *v) {
    const size_t n = (size_t)m->n_layer * 448 * m->d;
    memcpy(k, m->sk, sizeof(float) * n);
    memcpy(v, m->sv, sizeof(float) * n);
}
void oracle_dec_set_self_kv(oracle_dec *m, const float *k, const float *v) {
    const size_t n = (size_t)m->n_layer * 448 * m->d;
    memcpy(m->sk, k, sizeof(float) * n);
    memcpy(m->sv, v, sizeof(float) * n);
}

/* A.2 cross-KV: K_c = Wk enc (no bias), V_c = Wv enc + b, per decoder layer. enc: [1500][d] */
void oracle_dec_set_audio(oracle_dec *m, const float *enc) {
    const int d = m->d, T = 1500;
    float *e = (float *)malloc(sizeof(float) * (size_t)T * d);
    for (size_t i = 0; i < (size_t)T * d; i++) e[i] = rnd(m, enc[i]);
    for (int l = 0; l < m->n_layer; l++) {
        const float **w = m->lw + (size_t)l * W_COUNT;
        float *ck = m->ck + (size_t)l * T * d, *cv = m->cv + (size_t)l * T * d;
#pragma omp parallel for schedule(static)
        for (int t = 0; t < T; t++) {
            const float *x = e + (size_t)t * d;
            for (int j = 0; j < d; j++) {
                const float *wk = w[W_CKW] + (size_t)j * d, *wv = w[W_CVW] + (size_t)j * d;
                float a = 0.0f, b = 0.0f;
#pragma omp simd reduction(+ : a, b)
                for (int k = 0; k < d; k++) { a += wk[k] * x[k]; b += wv[k] * x[k]; }
                ck[(size_t)t * d + j] = rnd(m, a);
                cv[(size_t)t * d + j] = rnd(m, b + w[W_CVB][j]);
            }
        }
    }
    free(e);
}

/* A.3 one decoder step: token at position pos (self-KV rows [0, pos) must hold the previous tokens).
 * logits (may be NULL) receives n_vocab values.  aheads: n_aheads (layer, head) pairs; aprobs (may be NULL)
 * receives [n_aheads][1500] post-softmax cross-attention rows. */
int oracle_dec_step(oracle_dec *m, int token, int pos, float *logits, const int32_t *aheads, int n_aheads, float *aprobs) {
    const int d = m->d, H = m->n_head, dh = d / H, T = 1500;
    if (pos < 0 || pos >= 448 || token < 0 || token >= m->n_vocab) return -1;
    float *x = m->scratch, *h = x + d, *q = h + d, *att = q + d, *ff = att + d /* 4d */, *tmp = ff + 4 * d /* d */;
    float *sc = tmp + d; /* [H][1500] */
    const float scale = 1.0f / sqrtf((float)dh);
    for (int i = 0; i < d; i++) x[i] = m->tok_emb[(size_t)token * d + i] + m->pos[(size_t)pos * d + i];
    for (int l = 0; l < m->n_layer; l++) {
        const float **w = m->lw + (size_t)l * W_COUNT;
        float *sk = m->sk + (size_t)l * 448 * d, *sv = m->sv + (size_t)l * 448 * d;
        /* self attention */
        layer_norm(x, w[W_LN1G], w[W_LN1B], h, d);
        gemv(w[W_QW], w[W_QB], h, q, d, d);
        gemv(w[W_KW], NULL, h, sk + (size_t)pos * d, d, d);
        gemv(w[W_VW], w[W_VB], h, sv + (size_t)pos * d, d, d);
        if (m->bf16) /* storage mode: the library's self-KV cache is f16 (as whisper.cpp's kv_self is) */
            for (int i = 0; i < d; i++) {
                sk[(size_t)pos * d + i] = rnd16(m, sk[(size_t)pos * d + i]);
                sv[(size_t)pos * d + i] = rnd16(m, sv[(size_t)pos * d + i]);
            }
        for (int hh = 0; hh < H; hh++) {
            float *s = sc + (size_t)hh * T;
            float mx = -INFINITY;
            for (int t = 0; t <= pos; t++) {
                float a = 0.0f;
                for (int c = 0; c < dh; c++) a += q[hh * dh + c] * sk[(size_t)t * d + hh * dh + c];
                s[t] = a * scale;
                if (s[t] > mx) mx = s[t];
            }
            float sum = 0.0f;
            for (int t = 0; t <= pos; t++) { s[t] = expf(s[t] - mx); sum += s[t]; }
            const float inv = 1.0f / sum;
            for (int c = 0; c < dh; c++) {
                float a = 0.0f;
                for (int t = 0; t <= pos; t++) a += s[t] * inv * sv[(size_t)t * d + hh * dh + c];
                att[hh * dh + c] = a;
            }
        }
        gemv(w[W_OW], w[W_OB], att, tmp, d, d);
        for (int i = 0; i < d; i++) x[i] += tmp[i];
        /* cross attention */
        layer_norm(x, w[W_LN2G], w[W_LN2B], h, d);
        gemv(w[W_CQW], w[W_CQB], h, q, d, d);
        const float *ck = m->ck + (size_t)l * T * d, *cv = m->cv + (size_t)l * T * d;
#pragma omp parallel for schedule(static)
        for (int hh = 0; hh < H; hh++) {
            float *s = sc + (size_t)hh * T;
            float mx = -INFINITY;
            for (int t = 0; t < T; t++) {
                float a = 0.0f;
                for (int c = 0; c < dh; c++) a += q[hh * dh + c] * ck[(size_t)t * d + hh * dh + c];
                s[t] = a * scale;
                if (s[t] > mx) mx = s[t];
            }
            float sum = 0.0f;
            for (int t = 0; t < T; t++) { s[t] = expf(s[t] - mx); sum += s[t]; }
            const float inv = 1.0f / sum;
            for (int t = 0; t < T; t++) s[t] *= inv;
            for (int c = 0; c < dh; c++) {
                float a = 0.0f;
                for (int t = 0; t < T; t++) a += s[t] * cv[(size_t)t * d + hh * dh + c];
                att[hh * dh + c] = a;
            }
        }
        if (aprobs)
            for (int a = 0; a < n_aheads; a++)
                if (aheads[2 * a] == l) memcpy(aprobs + (size_t)a * T, sc + (size_t)aheads[2 * a + 1] * T, sizeof(float) * T);
        gemv(w[W_COW], w[W_COB], att, tmp, d, d);
        for (int i = 0; i < d; i++) x[i] += tmp[i];
        /* MLP */
        layer_norm(x, w[W_LN3G], w[W_LN3B], h, d);
        gemv(w[W_F1W], w[W_F1B], h, ff, 4 * d, d);
        for (int i = 0; i < 4 * d; i++) ff[i] = gelu_tanh_f(ff[i]);
        gemv(w[W_F2W], w[W_F2B], ff, tmp, d, 4 * d);
        for (int i = 0; i < d; i++) x[i] += tmp[i];
    }
    if (logits) {
        layer_norm(x, m->ln_g, m->ln_b, h, d);
        gemv(m->tok_emb, NULL, h, logits, m->n_vocab, d);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * A.4 whisper_process_logits + whisper_sample_token(best = true), temperature 0.
 * State of the (single) decoder: the tokens sampled so far in this window, has_ts, seek_delta.
 * logits is modified in place (masked); logprobs/probs are scratch of n_vocab floats.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    int suppress_blank, no_timestamps, suppress_nst;
    float max_initial_ts;
} oracle_logit_params;

static void compute_logprobs(const float *logits, int n, float *logprobs) {
    float mx = -INFINITY;
    for (int i = 0; i < n; i++) if (logits[i] > mx) mx = logits[i];
    float lse = 0.0f;
    for (int i = 0; i < n; i++) if (logits[i] > -INFINITY) lse += expf(logits[i] - mx);
    lse = logf(lse) + mx;
    for (int i = 0; i < n; i++) logprobs[i] = logits[i] > -INFINITY ? logits[i] - lse : -INFINITY;
}

float oracle_no_speech_prob(const float *logits, const oracle_vocab *v, float *scratch) {
    compute_logprobs(logits, v->n_vocab, scratch);
    return expf(scratch[v->nosp]);
}

void oracle_process_logits(float *logits, const oracle_vocab *v, const oracle_logit_params *lp, const oracle_token_data *tokens_cur,
                           int n_cur, int has_ts, int seek_delta, float *logprobs, float *probs) {
    const int n = v->n_vocab;
    const int is_initial = n_cur == 0;
    if (lp->suppress_blank && is_initial) {
        logits[v->eot] = -INFINITY;
        logits[v->space] = -INFINITY;
    }
    logits[v->not_] = -INFINITY;
    if (lp->no_timestamps)
        for (int i = v->beg; i < n; i++) logits[i] = -INFINITY;
    logits[v->sot] = -INFINITY;
    logits[v->nosp] = -INFINITY;
    logits[v->solm] = -INFINITY;
    logits[v->translate] = -INFINITY;
    logits[v->transcribe] = -INFINITY;
    logits[v->prev] = -INFINITY;
    for (int i = 0; i < v->n_langs; i++) logits[v->lang0 + i] = -INFINITY;
    /* timestamps have to appear in pairs, except directly before EOT */
    {
        const int last_was_ts = n_cur > 0 && tokens_cur[n_cur - 1].id >= v->beg;
        const int penult_was_ts = n_cur < 2 || tokens_cur[n_cur - 2].id >= v->beg;
        if (last_was_ts) {
            if (penult_was_ts) { for (int i = v->beg; i < n; i++) logits[i] = -INFINITY; }
            else { for (int i = 0; i < v->eot; i++) logits[i] = -INFINITY; }
        }
    }
    if (is_initial && lp->max_initial_ts > 0.0f) {
        const float precision = 30.0f / 1500.0f;
        const int tid0 = (int)roundf(lp->max_initial_ts / precision);
        for (int i = v->beg + tid0 + 1; i < n; i++) logits[i] = -INFINITY;
    }
    if (has_ts) {
        const int tid0 = seek_delta / 2;
        for (int i = v->beg; i < v->beg + tid0 && i < n; i++) logits[i] = -INFINITY;
    }
    compute_logprobs(logits, n, logprobs);
    {
        float ts_logprob = -INFINITY;
        float lmax = -INFINITY;
        for (int i = v->beg; i < n; i++) if (logprobs[i] > lmax) lmax = logprobs[i];
        float lse = 0.0f;
        for (int i = v->beg; i < n; i++) if (logprobs[i] > -INFINITY) lse += expf(logprobs[i] - lmax);
        if (lse > 0.0f) ts_logprob = logf(lse) + lmax;
        float max_text = -INFINITY;
        for (int i = 0; i < v->beg; i++) if (logprobs[i] > max_text) max_text = logprobs[i];
        if (ts_logprob > max_text)
            for (int i = 0; i < v->beg; i++) { logits[i] = -INFINITY; logprobs[i] = -INFINITY; }
    }
    for (int i = 0; i < n; i++) probs[i] = logits[i] == -INFINITY ? 0.0f : expf(logprobs[i]);
}

oracle_token_data oracle_sample_greedy(const float *probs, const float *logprobs, const oracle_vocab *v) {
    oracle_token_data r = {0, 0, 0.0f, 0.0f, 0.0f, 0.0f, -1, -1, -1, 0.0f};
    {
        double sum_ts = 0.0, max_ts = 0.0;
        for (int i = v->beg; i < v->n_vocab; i++) {
            sum_ts += probs[i];
            if (max_ts < probs[i]) { max_ts = probs[i]; r.tid = i; }
        }
        r.pt = (float)(max_ts / (sum_ts + 1e-10));
        r.ptsum = (float)sum_ts;
    }
    for (int i = 0; i < v->n_vocab; i++)
        if (r.p < probs[i]) { r.id = i; r.p = probs[i]; r.plog = logprobs[i]; }
    if (r.id >= v->beg) { r.tid = r.id; r.pt = r.p; }
    return r;
}

/* ------------------------------------------------------------------------------------------
 * A.4 one seek iteration of whisper_full_with_state for a single decoder at temperature 0
 * (greedy, no temperature fallback, no_context, no initial prompt beyond `prompt`).
 * Returns the number of tokens kept (result_len, or every sampled token when the decoder failed: whisper.cpp
 * resizes only non-failed sequences); tokens_out holds up to 224 entries.
 * info: [0] seek_delta  [1] failed  [2] completed  [3] n_sampled (before the resize)  [4] has_ts  [5] result_len
 * ------------------------------------------------------------------------------------------ */
int oracle_decode_window(oracle_dec *m, const oracle_vocab *v, const oracle_logit_params *lp, const int32_t *prompt, int n_prompt,
                         int seek, int seek_end, int single_segment, int delta_min, oracle_token_data *tokens_out, int32_t *info,
                         float *no_speech_prob, float *margins /* optional [224]: top1 - top2 filtered logit per step */) {
    const int n = v->n_vocab;
    float *logits = (float *)malloc(sizeof(float) * (size_t)n * 3);
    float *logprobs = logits + n, *probs = logprobs + n;
    for (int i = 0; i < n_prompt; i++) oracle_dec_step(m, prompt[i], i, i == n_prompt - 1 ? logits : NULL, NULL, 0, NULL);
    *no_speech_prob = oracle_no_speech_prob(logits, v, logprobs);
    int n_cur = 0, has_ts = 0, failed = 0, completed = 0, seek_delta = 3000, result_len = 0;
    const int n_max = 448 / 2 - 4;
    oracle_process_logits(logits, v, lp, tokens_out, n_cur, has_ts, seek_delta, logprobs, probs);
    for (int i = 0; i < n_max; i++) {
        oracle_token_data tok = oracle_sample_greedy(probs, logprobs, v);
        if (margins) {
            float top2 = -INFINITY;
            for (int k = 0; k < n; k++) if (k != tok.id && logits[k] > top2) top2 = logits[k];
            margins[i] = logits[tok.id] - top2;
        }
        tokens_out[n_cur++] = tok;
        if (tok.id > v->beg) {
            const int sd_new = 2 * (tok.id - v->beg);
            if (has_ts && seek_delta > sd_new && result_len < i) { failed = 1; break; }
            seek_delta = sd_new;
            result_len = i + 1;
            has_ts = 1;
        }
        if (tok.id == v->eot || (has_ts && seek + seek_delta + delta_min >= seek_end)) {
            if (result_len == 0 && !lp->no_timestamps) {
                if (seek + seek_delta + delta_min >= seek_end) result_len = i + 1;
                else { failed = 1; break; }
            }
            if (single_segment || lp->no_timestamps) { result_len = i + 1; seek_delta = 3000; }
            completed = 1;
            break;
        }
        if (i == n_max - 1 && (result_len == 0 || seek_delta < 3000 / 2)) { failed = 1; break; }
        oracle_dec_step(m, tok.id, n_prompt + i, logits, NULL, 0, NULL);
        oracle_process_logits(logits, v, lp, tokens_out, n_cur, has_ts, seek_delta, logprobs, probs);
    }
    info[0] = seek_delta; info[1] = failed; info[2] = completed; info[3] = n_cur; info[4] = has_ts; info[5] = result_len;
    free(logits);
    return failed ? n_cur : result_len;
}

/* ------------------------------------------------------------------------------------------
 * A.5 whisper_exp_compute_token_level_timestamps for one segment [t0, t1] (centiseconds).
 * vlen[j] = voice_length(token text) is supplied by the caller (depends on the vocabulary strings).
 * state3 = {t_beg, t_last, tid_last} carried across segments, as in whisper_state.
 * ------------------------------------------------------------------------------------------ */
static int ts_to_sample(int64_t t, int n_samples) {
    int64_t s = (t * 16000) / 100;
    if (s > n_samples - 1) s = n_samples - 1;
    if (s < 0) s = 0;
    return (int)s;
}
static int64_t sample_to_ts(int i) { return (100ll * i) / 16000; }

void oracle_token_timestamps(oracle_token_data *tokens, int n, int64_t t0, int64_t t1, const float *vlen, const float *energy,
                             int n_samples, int token_beg, int token_eot, float thold_pt, float thold_ptsum, int64_t *state3) {
    if (n_samples == 0 || n == 0) return;
    if (n == 1) { tokens[0].t0 = t0; tokens[0].t1 = t1; return; }
    int64_t *t_beg = &state3[0], *t_last = &state3[1], *tid_last = &state3[2];
    for (int j = 0; j < n; j++) {
        oracle_token_data *tk = &tokens[j];
        if (j == 0) {
            if (tk->id == token_beg) {
                tokens[j].t0 = t0; tokens[j].t1 = t0; tokens[j + 1].t0 = t0;
                *t_beg = t0; *t_last = t0; *tid_last = token_beg;
            } else {
                tokens[j].t0 = *t_last;
            }
        }
        const int64_t tt = *t_beg + 2 * (tk->tid - token_beg);
        tk->vlen = vlen[j];
        if (tk->pt > thold_pt && tk->ptsum > thold_ptsum && tk->tid > *tid_last && tt <= t1) {
            if (j > 0) tokens[j - 1].t1 = tt;
            tokens[j].t0 = tt;
            *tid_last = tk->tid;
        }
    }
    tokens[n - 2].t1 = t1; tokens[n - 1].t0 = t1; tokens[n - 1].t1 = t1;
    *t_last = t1;
    {
        int p0 = 0, p1 = 0;
        while (1) {
            while (p1 < n && tokens[p1].t1 < 0) p1++;
            if (p1 >= n) p1--;
            if (p1 > p0) {
                double psum = 0.0;
                for (int j = p0; j <= p1; j++) psum += tokens[j].vlen;
                const double dt = (double)(tokens[p1].t1 - tokens[p0].t0);
                for (int j = p0 + 1; j <= p1; j++) {
                    const double ct = tokens[j - 1].t0 + dt * tokens[j - 1].vlen / psum;
                    tokens[j - 1].t1 = (int64_t)ct;
                    tokens[j].t0 = (int64_t)ct;
                }
            }
            p1++;
            p0 = p1;
            if (p1 >= n) break;
        }
    }
    for (int j = 0; j < n - 1; j++) {
        if (tokens[j].t1 < 0) tokens[j + 1].t0 = tokens[j].t1;
        if (j > 0 && tokens[j - 1].t1 > tokens[j].t0) {
            tokens[j].t0 = tokens[j - 1].t1;
            tokens[j].t1 = tokens[j].t0 > tokens[j].t1 ? tokens[j].t0 : tokens[j].t1;
        }
    }
    {
        const int hw = 16000 / 8;
        for (int j = 0; j < n; j++) {
            if (tokens[j].id >= token_eot) continue;
            int s0 = ts_to_sample(tokens[j].t0, n_samples);
            int s1 = ts_to_sample(tokens[j].t1, n_samples);
            const int ss0 = s0 - hw > 0 ? s0 - hw : 0;
            const int ss1 = s1 + hw < n_samples ? s1 + hw : n_samples;
            const int ns = ss1 - ss0;
            float sum = 0.0f;
            for (int k = ss0; k < ss1; k++) sum += energy[k];
            const float thold = 0.5f * sum / ns;
            {
                int k = s0;
                if (energy[k] > thold && j > 0) {
                    while (k > 0 && energy[k] > thold) k--;
                    tokens[j].t0 = sample_to_ts(k);
                    if (tokens[j].t0 < tokens[j - 1].t1) tokens[j].t0 = tokens[j - 1].t1;
                    else s0 = k;
                } else {
                    while (energy[k] < thold && k < s1) k++;
                    s0 = k;
                    tokens[j].t0 = sample_to_ts(k);
                }
            }
            {
                int k = s1;
                if (energy[k] > thold) {
                    while (k < n_samples - 1 && energy[k] > thold) k++;
                    tokens[j].t1 = sample_to_ts(k);
                    if (j < ns - 1 && j + 1 < n && tokens[j].t1 > tokens[j + 1].t0) tokens[j].t1 = tokens[j + 1].t0;
                    else s1 = k;
                } else {
                    while (energy[k] < thold && k > s0) k--;
                    s1 = k;
                    tokens[j].t1 = sample_to_ts(k);
                }
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * A.6 steps 1-3: the teacher-forced decoder pass of whisper_exp_compute_token_level_timestamps_dtw.
 * seq: [sot, (lang), not, text..., eot] (n_seq tokens); aheads: n_aheads (layer, head) pairs.
 * w_out: [n_aheads][n_seq][n_audio] post-softmax cross attention restricted to the first n_audio positions.
 * ------------------------------------------------------------------------------------------ */
int oracle_dtw_attention(oracle_dec *m, const int32_t *seq, int n_seq, const int32_t *aheads, int n_aheads, int n_audio, float *w_out) {
    float *row = (float *)malloc(sizeof(float) * (size_t)n_aheads * 1500);
    for (int i = 0; i < n_seq; i++) {
        if (oracle_dec_step(m, seq[i], i, NULL, aheads, n_aheads, row) != 0) { free(row); return -1; }
        for (int a = 0; a < n_aheads; a++)
            memcpy(w_out + ((size_t)a * n_seq + i) * n_audio, row + (size_t)a * 1500, sizeof(float) * n_audio);
    }
    free(row);
    return 0;
}

/* A.6 step 8: stamp t_dtw on the window's text tokens from the DTW path. */
void oracle_dtw_stamp(oracle_token_data *tokens, int n_tokens, int token_eot, const int32_t *text_idx, const int32_t *time_idx,
                      int path_len, int seek) {
    int last_v = 0, ti = 0;
    for (int i = 0; i < path_len; i++) {
        const int v = text_idx[i];
        if (v != last_v) {
            const int64_t ts = (int64_t)time_idx[i] * 2 + seek;
            last_v = v;
            while (ti < n_tokens && !(tokens[ti].id < token_eot)) ti++;
            if (ti >= n_tokens) return;
            tokens[ti].t_dtw = ts;
            ti++;
        }
    }
}
