/*
 * wdr_oracle.c — CPU restatement of the numeric kernels on the hot path behind
 * Engine::transcribe_audio (reference call site: src/transcribe.rs:389 `state.full`,
 * src/transcribe.rs:466 `extractor.compute`).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked, imported or executed by the
 * product (whisper-diarize-rs_b200/, host/).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library.
 *
 * PARITY UNPINNED: the arithmetic of the reference lives in un-vendored third-party
 * dependencies (whisper-rs 0.15.0 -> whisper.cpp ~v1.7.6; pyannote-rs 0.3.1 -> knf-rs 0.3.1 /
 * kaldi-native-fbank; Cargo.lock:2498-2516, 1466-1475, 1075-1091) whose sources are absent
 * from /root/reference and the reference's own tests pin nothing on this path.
 * Each function restates the published upstream algorithm and is
 * cross-checked in tests/ against the independent OpenAI-lineage code in `transformers`
 * (mel, median filter, DTW) and `torchaudio` (Kaldi fbank).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#define WDR_SAMPLE_RATE 16000
#define WDR_N_FFT 400
#define WDR_HOP 160
#define WDR_CHUNK_S 30

/* ------------------------------------------------------------------------------------------
 * A.1 log-mel (whisper.cpp log_mel_spectrogram; driven from src/transcribe.rs:389).
 * fp32 recursive radix-2 FFT with a direct DFT for odd sizes (400->200->100->50->25->DFT),
 * sin/cos tables of 400 float entries, periodic Hann, mel accumulation in double.
 * ------------------------------------------------------------------------------------------ */
static float g_sin[WDR_N_FFT], g_cos[WDR_N_FFT], g_hann[WDR_N_FFT];
static int g_tables_ready = 0;

static void fill_tables(void) {
    if (g_tables_ready) return;
    for (int i = 0; i < WDR_N_FFT; i++) {
        double theta = (2.0 * M_PI * i) / WDR_N_FFT;
        g_sin[i] = sinf((float)theta);
        g_cos[i] = cosf((float)theta);
        g_hann[i] = (float)(0.5 * (1.0 - cosf((float)((2.0 * M_PI * i) / WDR_N_FFT))));
    }
    g_tables_ready = 1;
}

static void dft_f32(const float *in, int N, float *out) {
    const int step = WDR_N_FFT / N;
    for (int k = 0; k < N; k++) {
        float re = 0, im = 0;
        for (int n = 0; n < N; n++) {
            int idx = (k * n * step) % WDR_N_FFT;
            re += in[n] * g_cos[idx];
            im -= in[n] * g_sin[idx];
        }
        out[2 * k] = re;
        out[2 * k + 1] = im;
    }
}

/* in: N real values followed by scratch (>= N more); out: 2N values followed by scratch. */
static void fft_f32(float *in, int N, float *out) {
    if (N == 1) { out[0] = in[0]; out[1] = 0; return; }
    const int half = N / 2;
    if (N - half * 2 == 1) { dft_f32(in, N, out); return; }
    float *even = in + N;
    for (int i = 0; i < half; i++) even[i] = in[2 * i];
    float *even_fft = out + 2 * N;
    fft_f32(even, half, even_fft);
    float *odd = even;
    for (int i = 0; i < half; i++) odd[i] = in[2 * i + 1];
    float *odd_fft = even_fft + N;
    fft_f32(odd, half, odd_fft);
    const int step = WDR_N_FFT / N;
    for (int k = 0; k < half; k++) {
        int idx = k * step;
        float re = g_cos[idx], im = -g_sin[idx];
        float ro = odd_fft[2 * k], io = odd_fft[2 * k + 1];
        out[2 * k] = even_fft[2 * k] + re * ro - im * io;
        out[2 * k + 1] = even_fft[2 * k + 1] + re * io + im * ro;
        out[2 * (k + half)] = even_fft[2 * k] - re * ro + im * io;
        out[2 * (k + half) + 1] = even_fft[2 * k + 1] - re * io - im * ro;
    }
}

/* n_len of the padded buffer: (n + 30 s + 2*200 - 400) / 160 */
int oracle_mel_n_len(int n_samples) {
    return (n_samples + WDR_SAMPLE_RATE * WDR_CHUNK_S + WDR_N_FFT - WDR_N_FFT) / WDR_HOP;
}

/* filters: [n_mel][201] fp32.  out: [n_mel][n_len] fp32 mel-major.  normalize: 1 = clamp to
 * global max-8 and (x+4)/4 over the whole buffer (whisper.cpp behaviour); 0 = raw log10.
 * Returns n_len. */
int oracle_log_mel(const float *pcm, int n, const float *filters, int n_mel, int normalize,
                   float *out) {
    fill_tables();
    const int n_bins = 1 + WDR_N_FFT / 2;
    const int64_t pad1 = WDR_SAMPLE_RATE * WDR_CHUNK_S, pad2 = WDR_N_FFT / 2;
    const int64_t np = n + pad1 + 2 * pad2;
    float *x = (float *)calloc((size_t)np, sizeof(float));
    memcpy(x + pad2, pcm, (size_t)n * sizeof(float));
    /* reflective pad 200 at the beginning: reverse_copy(samples+1, samples+1+200) */
    for (int i = 0; i < pad2; i++) x[i] = (1 + (pad2 - 1 - i)) < n ? pcm[1 + (pad2 - 1 - i)] : 0.0f;
    const int n_len = (int)((np - WDR_N_FFT) / WDR_HOP);
    const int64_t n_frames_data = np / WDR_HOP + 1;
    const int n_calc = (int)(n_frames_data < n_len ? n_frames_data : n_len);
#pragma omp parallel
    {
        float fft_in[2 * WDR_N_FFT + 16];
        float fft_out[8 * WDR_N_FFT + 16];
#pragma omp for schedule(static)
        for (int i = 0; i < n_calc; i++) {
            const int64_t off = (int64_t)i * WDR_HOP;
            int64_t lim = np - off; if (lim > WDR_N_FFT) lim = WDR_N_FFT;
            for (int j = 0; j < lim; j++) fft_in[j] = g_hann[j] * x[off + j];
            for (int j = (int)(lim < 0 ? 0 : lim); j < WDR_N_FFT; j++) fft_in[j] = 0.0f;
            fft_f32(fft_in, WDR_N_FFT, fft_out);
            for (int j = 0; j < n_bins; j++)
                fft_out[j] = fft_out[2 * j] * fft_out[2 * j] + fft_out[2 * j + 1] * fft_out[2 * j + 1];
            for (int m = 0; m < n_mel; m++) {
                double sum = 0.0;
                const float *f = filters + (size_t)m * n_bins;
                for (int k = 0; k < n_bins; k++) sum += (double)(fft_out[k] * f[k]);
                if (sum < 1e-10) sum = 1e-10;
                out[(size_t)m * n_len + i] = (float)log10(sum);
            }
        }
    }
    const float floor_v = (float)log10(1e-10);
    for (int i = n_calc; i < n_len; i++)
        for (int m = 0; m < n_mel; m++) out[(size_t)m * n_len + i] = floor_v;
    if (normalize) {
        double mmax = -1e20;
        const size_t tot = (size_t)n_mel * n_len;
        for (size_t i = 0; i < tot; i++) if (out[i] > mmax) mmax = out[i];
        mmax -= 8.0;
        for (size_t i = 0; i < tot; i++) {
            double v = out[i];
            if (v < mmax) v = mmax;
            out[i] = (float)((v + 4.0) / 4.0);
        }
    }
    free(x);
    return n_len;
}

/* ------------------------------------------------------------------------------------------
 * A.6 median filter (whisper.cpp median_filter custom op; width 7, reflect indexing),
 * applied along the last (audio) axis of w[H][N][M].
 * ------------------------------------------------------------------------------------------ */
static int cmp_f32(const void *a, const void *b) {
    float x = *(const float *)a, y = *(const float *)b;
    return (x > y) - (x < y);
}

int oracle_median_filter(const float *w, int H, int N, int M, int width, float *out) {
    if (width <= 0 || (width & 1) == 0 || width > 63) return -1;
    const int hw = width / 2;
    if (M <= hw) return -2; /* reflect index would leave the row */
    for (int64_t r = 0; r < (int64_t)H * N; r++) {
        const float *src = w + r * M;
        float *dst = out + r * M;
        float win[64];
        for (int j = 0; j < M; j++) {
            for (int k = 0; k < width; k++) {
                int idx = j + k - hw;
                if (idx < 0) idx = -idx;
                else if (idx >= M) idx = 2 * (M - 1) - idx;
                win[k] = src[idx];
            }
            qsort(win, (size_t)width, sizeof(float), cmp_f32);
            dst[j] = win[hw];
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * A.6 dtw_and_backtrace (whisper.cpp; same 3-way tie-break as OpenAI's dtw_cpu):
 * x[N][M] fp32 cost; outputs the path (text_idx, time_idx), length returned in *path_len
 * (<= N+M).  Also optionally returns the cost/trace matrices ((N+1)*(M+1)) for kernel checks.
 * ------------------------------------------------------------------------------------------ */
int oracle_dtw(const float *x, int N, int M, int32_t *text_idx, int32_t *time_idx, int *path_len,
               float *cost_out, int32_t *trace_out) {
    if (N <= 0 || M <= 0) { *path_len = 0; return 0; }
    const size_t W = (size_t)M + 1;
    float *cost = (float *)malloc((size_t)(N + 1) * W * sizeof(float));
    int32_t *trace = (int32_t *)malloc((size_t)(N + 1) * W * sizeof(int32_t));
    for (size_t i = 0; i < (size_t)(N + 1) * W; i++) { cost[i] = INFINITY; trace[i] = -1; }
    cost[0] = 0.0f;
    for (int j = 1; j <= M; j++) {
        for (int i = 1; i <= N; i++) {
            float c0 = cost[(size_t)(i - 1) * W + (j - 1)];
            float c1 = cost[(size_t)(i - 1) * W + j];
            float c2 = cost[(size_t)i * W + (j - 1)];
            float c; int32_t t;
            if (c0 < c1 && c0 < c2) { c = c0; t = 0; }
            else if (c1 < c0 && c1 < c2) { c = c1; t = 1; }
            else { c = c2; t = 2; }
            cost[(size_t)i * W + j] = x[(size_t)(i - 1) * M + (j - 1)] + c;
            trace[(size_t)i * W + j] = t;
        }
    }
    if (cost_out) memcpy(cost_out, cost, (size_t)(N + 1) * W * sizeof(float));
    for (int j = 0; j <= M; j++) trace[j] = 2;
    for (int i = 0; i <= N; i++) trace[(size_t)i * W] = 1;
    if (trace_out) memcpy(trace_out, trace, (size_t)(N + 1) * W * sizeof(int32_t));
    int i = N, j = M, n = 0;
    int32_t *ti = (int32_t *)malloc((size_t)(N + M + 2) * sizeof(int32_t));
    int32_t *tj = (int32_t *)malloc((size_t)(N + M + 2) * sizeof(int32_t));
    int rc = 0;
    while (i > 0 || j > 0) {
        ti[n] = i - 1; tj[n] = j - 1; n++;
        int32_t t = trace[(size_t)i * W + j];
        if (t == 0) { i--; j--; }
        else if (t == 1) { i--; }
        else if (t == 2) { j--; }
        else { rc = -1; break; }
    }
    for (int k = 0; k < n; k++) { text_idx[k] = ti[n - 1 - k]; time_idx[k] = tj[n - 1 - k]; }
    *path_len = n;
    free(ti); free(tj); free(cost); free(trace);
    return rc;
}

/* ------------------------------------------------------------------------------------------
 * A.6 steps 3-6: alignment-head weights [n_tokens][n_audio][H] -> cost x[N][M]
 *   ggml_norm over the token axis (eps 1e-9, population variance), median filter width 7
 *   over audio, mean over heads, *(-1), drop sot_len rows and the final ([EOT]) row.
 * w layout here: [H][n_tokens][n_audio] (head-major, as the device kernel stores it).
 * ------------------------------------------------------------------------------------------ */
int oracle_dtw_cost(const float *w, int H, int n_tokens, int n_audio, int sot_len, int width,
                    float *x_out /* [(n_tokens-sot_len-1)][n_audio] */) {
    const size_t plane = (size_t)n_tokens * n_audio;
    float *nrm = (float *)malloc((size_t)H * plane * sizeof(float));
    float *med = (float *)malloc((size_t)H * plane * sizeof(float));
    for (int h = 0; h < H; h++)
        for (int a = 0; a < n_audio; a++) {
            /* ggml_norm: mean and variance accumulated in double (ggml_float), y=(x-mean)*1/sqrt(var+eps) */
            double sum = 0.0;
            for (int t = 0; t < n_tokens; t++) sum += (double)w[h * plane + (size_t)t * n_audio + a];
            float mean = (float)(sum / n_tokens);
            double sum2 = 0.0;
            for (int t = 0; t < n_tokens; t++) {
                float v = w[h * plane + (size_t)t * n_audio + a] - mean;
                nrm[h * plane + (size_t)t * n_audio + a] = v;
                sum2 += (double)(v * v);
            }
            float variance = (float)(sum2 / n_tokens);
            const float scale = 1.0f / sqrtf(variance + 1e-9f);
            for (int t = 0; t < n_tokens; t++) nrm[h * plane + (size_t)t * n_audio + a] *= scale;
        }
    int rc = oracle_median_filter(nrm, H, n_tokens, n_audio, width, med);
    const int N = n_tokens - sot_len - 1;
    if (rc == 0)
        for (int t = 0; t < N; t++)
            for (int a = 0; a < n_audio; a++) {
                float s = 0.0f;
                for (int h = 0; h < H; h++) s += med[h * plane + (size_t)(t + sot_len) * n_audio + a];
                x_out[(size_t)t * n_audio + a] = -(s / (float)H);
            }
    free(nrm); free(med);
    return rc;
}

/* ------------------------------------------------------------------------------------------
 * A.9 Kaldi fbank (kaldi-native-fbank via knf-rs compute_fbank; call site
 * src/transcribe.rs:466): 25 ms / 10 ms, dither 0, snip_edges, remove_dc_offset, preemph 0.97,
 * povey window, 512-pt FFT power spectrum, 80 HTK-mel triangular bins 20..8000 Hz,
 * log(max(e, FLT_EPSILON)); then per-column mean subtraction over frames (pyannote-rs).
 * in: raw int16-scale samples as fp32.  out [T][n_bins].  Returns T (0 if too short).
 * ------------------------------------------------------------------------------------------ */
/* kaldi-native-fbank MelScale: float arithmetic (mel-computations.h) */
static inline float mel_htk(float f) { return 1127.0f * logf(1.0f + f / 700.0f); }

int oracle_fbank_frames(int n) { return n < 400 ? 0 : 1 + (n - 400) / 160; }

int oracle_kaldi_fbank(const float *wave, int n, int n_bins, int subtract_mean, float *out) {
    const int flen = 400, shift = 160, nfft = 512, nb = nfft / 2;
    const int T = oracle_fbank_frames(n);
    if (T == 0) return 0;
    /* povey window */
    float win[400];
    for (int i = 0; i < flen; i++) {
        double a = 2.0 * M_PI / (flen - 1);
        win[i] = (float)pow(0.5 - 0.5 * cos(a * i), 0.85);
    }
    /* mel banks (kaldi MelBanks): low 20, high nyquist, no vtln */
    const float nyq = 8000.0f, low = 20.0f, high = nyq;
    const float fft_bin_w = (float)WDR_SAMPLE_RATE / nfft;
    const float mlow = mel_htk(low), mhigh = mel_htk(high);
    const float mdelta = (mhigh - mlow) / (n_bins + 1);
    float *bank = (float *)calloc((size_t)n_bins * nb, sizeof(float));
    for (int b = 0; b < n_bins; b++) {
        float lm = mlow + b * mdelta, cm = mlow + (b + 1) * mdelta, rm = mlow + (b + 2) * mdelta;
        for (int i = 0; i < nb; i++) {
            float mel = mel_htk(fft_bin_w * i);
            if (mel > lm && mel < rm) {
                float wgt = mel <= cm ? (mel - lm) / (cm - lm) : (rm - mel) / (rm - cm);
                bank[(size_t)b * nb + i] = wgt;
            }
        }
    }
#pragma omp parallel
    {
        float fr[512];
        double re[257], im[257];
#pragma omp for schedule(static)
        for (int t = 0; t < T; t++) {
            const float *src = wave + (size_t)t * shift;
            float sum = 0.0f;
            for (int i = 0; i < flen; i++) sum += src[i];
            float mean = sum / flen;
            for (int i = 0; i < flen; i++) fr[i] = src[i] - mean;
            for (int i = flen - 1; i > 0; i--) fr[i] -= 0.97f * fr[i - 1];
            fr[0] -= 0.97f * fr[0];
            for (int i = 0; i < flen; i++) fr[i] *= win[i];
            for (int i = flen; i < nfft; i++) fr[i] = 0.0f;
            /* power spectrum via direct real DFT in double (oracle: clarity over speed) */
            for (int k = 0; k <= nb; k++) { re[k] = 0; im[k] = 0; }
            for (int k = 0; k < nb; k++) {
                double sr = 0, si = 0;
                for (int i = 0; i < flen; i++) {
                    double ang = -2.0 * M_PI * (double)((k * i) % nfft) / nfft;
                    sr += fr[i] * cos(ang); si += fr[i] * sin(ang);
                }
                re[k] = sr; im[k] = si;
            }
            for (int b = 0; b < n_bins; b++) {
                float e = 0.0f;
                for (int k = 0; k < nb; k++) {
                    float p = (float)(re[k] * re[k] + im[k] * im[k]);
                    e += bank[(size_t)b * nb + k] * p;
                }
                if (e < FLT_EPSILON) e = FLT_EPSILON;
                out[(size_t)t * n_bins + b] = logf(e);
            }
        }
    }
    if (subtract_mean)
        for (int b = 0; b < n_bins; b++) {
            float s = 0.0f;
            for (int t = 0; t < T; t++) s += out[(size_t)t * n_bins + b];
            s /= T;
            for (int t = 0; t < T; t++) out[(size_t)t * n_bins + b] -= s;
        }
    free(bank);
    return T;
}

/* ------------------------------------------------------------------------------------------
 * A.4/A.5 get_signal_energy (whisper.cpp; n_samples_per_half_window = 32): moving average of
 * |x| over [i-hw, i+hw] clipped to the buffer, always normalised by 2*hw+1.
 * ------------------------------------------------------------------------------------------ */
void oracle_signal_energy(const float *signal, int n, int hw, float *out) {
    for (int i = 0; i < n; i++) {
        float sum = 0;
        for (int j = -hw; j <= hw; j++)
            if (i + j >= 0 && i + j < n) sum += fabsf(signal[i + j]);
        out[i] = sum / (2 * hw + 1);
    }
}

/* ==========================================================================================
 * A.2 Whisper encoder (whisper.cpp whisper_encode_internal): conv1d(k3,s1,p1)+GELU ->
 * conv1d(k3,s2,p1)+GELU -> + sinusoidal positions -> n_layer x {LN, MHA (key has no bias),
 * +res, LN, MLP(4x, GELU), +res} -> ln_post.  fp32 throughout; GELU = ggml's tanh form
 * (the f16 lookup table ggml uses on CPU is not reproduced; the 1e-2 tolerance absorbs it).
 * Weights arrive as one flat fp32 array in the order documented in oracle/weights.py.
 * ========================================================================================== */
static inline float gelu_tanh_f(float x) {
    return 0.5f * x * (1.0f + tanhf(0.79788456080286535587989211986876f * x * (1.0f + 0.044715f * x * x)));
}

/* C[M][N] (ldc) = A[M][K] (lda) * W[N][K]^T (ldw) + bias[N] (bias may be NULL); accumulate==1 adds into C. */
static void gemm_nt(const float *A, int lda, const float *W, int ldw, const float *bias, float *C, int ldc,
                    int M, int N, int K, int accumulate) {
    /* 4 x 4 register tile of dot products along K (both operands are K-contiguous): 8 vector loads feed 16 multiply-adds */
#pragma omp parallel for schedule(static)
    for (int i0 = 0; i0 < M; i0 += 4) {
        const int im = (M - i0) < 4 ? (M - i0) : 4;
        for (int j0 = 0; j0 < N; j0 += 4) {
            const int jm = (N - j0) < 4 ? (N - j0) : 4;
            float acc[4][4] = {{0}};
            if (im == 4 && jm == 4) {
                const float *a0 = A + (size_t)i0 * lda, *a1 = a0 + lda, *a2 = a1 + lda, *a3 = a2 + lda;
                const float *w0 = W + (size_t)j0 * ldw, *w1 = w0 + ldw, *w2 = w1 + ldw, *w3 = w2 + ldw;
                float s00 = 0, s01 = 0, s02 = 0, s03 = 0, s10 = 0, s11 = 0, s12 = 0, s13 = 0;
                float s20 = 0, s21 = 0, s22 = 0, s23 = 0, s30 = 0, s31 = 0, s32 = 0, s33 = 0;
#pragma omp simd reduction(+ : s00, s01, s02, s03, s10, s11, s12, s13, s20, s21, s22, s23, s30, s31, s32, s33)
                for (int k = 0; k < K; k++) {
                    const float x0 = a0[k], x1 = a1[k], x2 = a2[k], x3 = a3[k];
                    const float y0 = w0[k], y1 = w1[k], y2 = w2[k], y3 = w3[k];
                    s00 += x0 * y0; s01 += x0 * y1; s02 += x0 * y2; s03 += x0 * y3;
                    s10 += x1 * y0; s11 += x1 * y1; s12 += x1 * y2; s13 += x1 * y3;
                    s20 += x2 * y0; s21 += x2 * y1; s22 += x2 * y2; s23 += x2 * y3;
                    s30 += x3 * y0; s31 += x3 * y1; s32 += x3 * y2; s33 += x3 * y3;
                }
                acc[0][0] = s00; acc[0][1] = s01; acc[0][2] = s02; acc[0][3] = s03;
                acc[1][0] = s10; acc[1][1] = s11; acc[1][2] = s12; acc[1][3] = s13;
                acc[2][0] = s20; acc[2][1] = s21; acc[2][2] = s22; acc[2][3] = s23;
                acc[3][0] = s30; acc[3][1] = s31; acc[3][2] = s32; acc[3][3] = s33;
            } else {
                for (int ii = 0; ii < im; ii++)
                    for (int jj = 0; jj < jm; jj++) {
                        const float *a = A + (size_t)(i0 + ii) * lda, *w = W + (size_t)(j0 + jj) * ldw;
                        float sum = 0.0f;
#pragma omp simd reduction(+ : sum)
                        for (int k = 0; k < K; k++) sum += a[k] * w[k];
                        acc[ii][jj] = sum;
                    }
            }
            for (int ii = 0; ii < im; ii++)
                for (int jj = 0; jj < jm; jj++) {
                    const float v = acc[ii][jj] + (bias ? bias[j0 + jj] : 0.0f);
                    float *c = C + (size_t)(i0 + ii) * ldc + j0 + jj;
                    *c = accumulate ? *c + v : v;
                }
        }
    }
}

static void layer_norm_rows(const float *x, const float *g, const float *b, float *y, int rows, int d) {
#pragma omp parallel for schedule(static)
    for (int r = 0; r < rows; r++) {
        const float *xr = x + (size_t)r * d;
        float *yr = y + (size_t)r * d;
        double s = 0.0;
        for (int i = 0; i < d; i++) s += xr[i];
        const float mean = (float)(s / d);
        double q = 0.0;
        for (int i = 0; i < d; i++) { const float v = xr[i] - mean; q += (double)v * v; }
        const float rstd = 1.0f / sqrtf((float)(q / d) + 1e-5f);
        for (int i = 0; i < d; i++) yr[i] = (xr[i] - mean) * rstd * g[i] + b[i];
    }
}

/* softmax(Q K^T * scale) V for one head; q,k,v,o are [T][ld] slices. */
static void attention_head(const float *q, const float *k, const float *v, float *o, int Tq, int Tk, int dh, int ld, int ldo,
                           float scale, float *probs_out /* optional [Tq][Tk] */) {
#pragma omp parallel
    {
        float *s = (float *)malloc(sizeof(float) * (size_t)Tk);
#pragma omp for schedule(static)
        for (int i = 0; i < Tq; i++) {
            const float *qi = q + (size_t)i * ld;
            float mx = -INFINITY;
            for (int j = 0; j < Tk; j++) {
                const float *kj = k + (size_t)j * ld;
                float d = 0.0f;
#pragma omp simd reduction(+ : d)
                for (int c = 0; c < dh; c++) d += qi[c] * kj[c];
                s[j] = d * scale;
                if (s[j] > mx) mx = s[j];
            }
            float sum = 0.0f;
            for (int j = 0; j < Tk; j++) { s[j] = expf(s[j] - mx); sum += s[j]; }
            const float inv = 1.0f / sum;
            float *oi = o + (size_t)i * ldo;
            for (int c = 0; c < dh; c++) oi[c] = 0.0f;
            for (int j = 0; j < Tk; j++) {
                const float p = s[j] * inv;
                if (probs_out) probs_out[(size_t)i * Tk + j] = p;
                const float *vj = v + (size_t)j * ld;
#pragma omp simd
                for (int c = 0; c < dh; c++) oi[c] += p * vj[c];
            }
        }
        free(s);
    }
}

/* mel: [n_mel][3000] normalised.  wts: flat weights (oracle/weights.py: pack_encoder).  out: [1500][d]. */
int oracle_whisper_encode(const float *mel, int n_mel, int d, int n_head, int n_layer, const float *wts, float *out) {
    const int T0 = 3000, T = 1500, dh = d / n_head;
    const float *p = wts;
    const float *c1w = p; p += (size_t)d * n_mel * 3;
    const float *c1b = p; p += d;
    const float *c2w = p; p += (size_t)d * d * 3;
    const float *c2b = p; p += d;
    const float *pos = p; p += (size_t)T * d;
    float *a1 = (float *)calloc((size_t)T0 * d, sizeof(float));
    float *x = (float *)calloc((size_t)T * d, sizeof(float));
    float *h = (float *)malloc(sizeof(float) * (size_t)T * d);
    float *qkv = (float *)malloc(sizeof(float) * (size_t)T * 3 * d);
    float *att = (float *)malloc(sizeof(float) * (size_t)T * d);
    float *ff = (float *)malloc(sizeof(float) * (size_t)T * 4 * d);
    /* conv1: out[t][o] = gelu(b[o] + sum_c sum_k w[o][c][k] * mel[c][t + k - 1]) */
#pragma omp parallel for schedule(static)
    for (int t = 0; t < T0; t++)
        for (int o = 0; o < d; o++) {
            float s = c1b[o];
            for (int c = 0; c < n_mel; c++)
                for (int k = 0; k < 3; k++) {
                    const int tt = t + k - 1;
                    if (tt >= 0 && tt < T0) s += c1w[((size_t)o * n_mel + c) * 3 + k] * mel[(size_t)c * T0 + tt];
                }
            a1[(size_t)t * d + o] = gelu_tanh_f(s);
        }
    /* conv2 (stride 2): out[j][o] = gelu(b[o] + sum_c sum_k w[o][c][k] * a1[2j + k - 1][c]) + pos[j][o] */
#pragma omp parallel for schedule(static)
    for (int j = 0; j < T; j++)
        for (int o = 0; o < d; o++) {
            float s = c2b[o];
            for (int k = 0; k < 3; k++) {
                const int tt = 2 * j + k - 1;
                if (tt < 0 || tt >= T0) continue;
                const float *ar = a1 + (size_t)tt * d;
                const float *wr = c2w + (size_t)o * d * 3 + k;
                float acc = 0.0f;
                for (int c = 0; c < d; c++) acc += wr[(size_t)c * 3] * ar[c];
                s += acc;
            }
            x[(size_t)j * d + o] = gelu_tanh_f(s) + pos[(size_t)j * d + o];
        }
    const float scale = 1.0f / sqrtf((float)dh);
    for (int l = 0; l < n_layer; l++) {
        const float *ln1g = p; p += d;
        const float *ln1b = p; p += d;
        const float *qw = p; p += (size_t)d * d;
        const float *qb = p; p += d;
        const float *kw = p; p += (size_t)d * d;
        const float *vw = p; p += (size_t)d * d;
        const float *vb = p; p += d;
        const float *ow = p; p += (size_t)d * d;
        const float *ob = p; p += d;
        const float *ln2g = p; p += d;
        const float *ln2b = p; p += d;
        const float *f1w = p; p += (size_t)4 * d * d;
        const float *f1b = p; p += 4 * d;
        const float *f2w = p; p += (size_t)4 * d * d;
        const float *f2b = p; p += d;
        layer_norm_rows(x, ln1g, ln1b, h, T, d);
        gemm_nt(h, d, qw, d, qb, qkv, 3 * d, T, d, d, 0);
        gemm_nt(h, d, kw, d, NULL, qkv + d, 3 * d, T, d, d, 0);
        gemm_nt(h, d, vw, d, vb, qkv + 2 * d, 3 * d, T, d, d, 0);
        for (int hh = 0; hh < n_head; hh++)
            attention_head(qkv + hh * dh, qkv + d + hh * dh, qkv + 2 * d + hh * dh, att + hh * dh, T, T, dh, 3 * d, d, scale, NULL);
        gemm_nt(att, d, ow, d, ob, x, d, T, d, d, 1);
        layer_norm_rows(x, ln2g, ln2b, h, T, d);
        gemm_nt(h, d, f1w, d, f1b, ff, 4 * d, T, 4 * d, d, 0);
        for (size_t i = 0; i < (size_t)T * 4 * d; i++) ff[i] = gelu_tanh_f(ff[i]);
        gemm_nt(ff, 4 * d, f2w, 4 * d, f2b, x, d, T, d, 4 * d, 1);
    }
    layer_norm_rows(x, p, p + d, out, T, d);
    free(a1); free(x); free(h); free(qkv); free(att); free(ff);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Seeded weight synthesis (oracle/weights.py documents the generator): dst[i] = offset + u_i * scale with
 * u_i = (int(splitmix64(key + i) >> 40) - 2^23) / 2^23, optionally rounded to bf16 (round-to-nearest-even).
 * Same arithmetic as the numpy path in weights.py, in C so that large-v3 (1.5 G parameters) is generated in seconds.
 * ------------------------------------------------------------------------------------------ */
static inline uint64_t splitmix64_c(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
void oracle_synth_fill(float *dst, int64_t n, uint64_t key, float offset, float scale, int bf16) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        const uint64_t z = splitmix64_c(key + (uint64_t)i);
        const int k = (int)(z >> 40);
        const float u = (float)(k - 8388608) * (1.0f / 8388608.0f);
        volatile float prod = u * scale; /* no FMA contraction: matches numpy's separate multiply and add */
        float v = offset + prod;
        if (bf16) {
            uint32_t b;
            memcpy(&b, &v, 4);
            b = (b + 0x7fffu + ((b >> 16) & 1u)) & 0xffff0000u;
            memcpy(&v, &b, 4);
        }
        dst[i] = v;
    }
}


/* ------------------------------------------------------------------------------------------------
 * OpenMP thread control for the timed CPU arm (bench.py): a launcher may export OMP_NUM_THREADS=1
 * (torch.distributed.run does for multi-rank launches), which would silently run the port on one
 * thread while the bench line claims all cores.  The bench sets the count explicitly and reports
 * what the runtime really uses.
 * ------------------------------------------------------------------------------------------------ */
#ifdef _OPENMP
#include <omp.h>
void oracle_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int oracle_max_threads(void) { return omp_get_max_threads(); }
int oracle_threads_in_parallel(void) {
    int n = 1;
#pragma omp parallel
    {
#pragma omp single
        n = omp_get_num_threads();
    }
    return n;
}
#else
void oracle_set_threads(int n) { (void)n; }
int oracle_max_threads(void) { return 1; }
int oracle_threads_in_parallel(void) { return 1; }
#endif
