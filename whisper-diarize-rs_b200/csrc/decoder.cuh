// decoder.cuh — KV-cached Whisper decoder over a batch of windows: workspace and entry points (decoder.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>
#include "model.cuh"
#include "profile.cuh"
#include "vocab.cuh"

namespace wdr {

// A decode batch has WINDOWS (each owns a cross-KV cache: 245.8 MB for large-v3, so their number is bounded by memory) and ROWS
// (decoders: greedy = one per window; beam search / best_of = several per window, sharing its cross cache).  Greedy batches are one
// 128-row M tile of the tcgen05 GEMM; beam batches run up to kDecMaxRows rows (five M tiles) so that the weight-streaming chain is
// paid once per iteration for all beams of up to 128 windows instead of once per 25 windows.
constexpr int kDecMaxWindows = 128;
constexpr int kDecMaxRows = 640;
constexpr int kDecMaxBatch = kDecMaxRows;  // row capacity of the per-row beam tables
constexpr int kDecMaxTokens = 224;  // sampled tokens kept per window (n_text_ctx/2)
constexpr int kDecSeqCap = WDR_TEXT_CTX;

// per-window decoder state (whisper_decoder + whisper_sequence fields the loop needs), device resident
struct DecWinState {
    int32_t n_cur, has_ts, seek_delta, result_len, failed, completed, seek, seek_end;
    float no_speech_prob;
    int32_t pad[3];
};

struct SampleParams {
    int n_vocab, eot, sot, translate, transcribe, solm, prev, nosp, not_, beg, lang0, n_langs, space;
    int suppress_blank, no_timestamps, single_segment, delta_min, n_max;
    int initial_tid0;  // round(max_initial_ts / 0.02), or -1 when max_initial_ts <= 0
};

// ---- beam search (rows = windows x beams) ----
constexpr int kBeamMax = 8;
struct BeamRow {   // what whisper_process_logits needs to know about a beam's history
    int32_t active, n_cur, last_id, penult_id, has_ts, seek_delta;
};
struct BeamCand {  // one entry of whisper_sample_token_topk
    int32_t id, tid;
    float p, plog, pt, ptsum;
};

struct DecoderWorkspace {
    int cap_B = 0;
    int d = 0, n_layer = 0, n_head = 0;
    int64_t ldv = 0;                   // logits row stride (n_vocab rounded up to 8)
    __nv_bfloat16* enc_bf16 = nullptr; // [B][1500][d]   ln_post output, the cross-KV GEMM's A operand
    std::vector<__nv_bfloat16*> ckv;   // per layer [B][H][K | V][1500][64] bf16: head-major, each (window, head) block contiguous
    __half* sk = nullptr;              // [L][B][H] blocks of 448 x 64 f16 (whisper.cpp: kv_self is f16): keys transposed in blocks of 32 positions
    __half* sv = nullptr;              //            values row-major [448][64]
    float* x = nullptr;                // [B][d]  residual stream
    __nv_bfloat16* h = nullptr;        // [2][B][d]   (hi, lo) planes of the LayerNorm output
    __nv_bfloat16* att = nullptr;      // [2][B][d]
    __nv_bfloat16* ff = nullptr;       // [2][B][4d]
    float* part = nullptr;             // split-K partial sums
    size_t part_elems = 0;
    float* logits = nullptr;           // [B][ldv]
    int32_t* seq = nullptr;            // [B][448] tokens fed to the decoder (prompt, then sampled / teacher-forced)
    wdr_token_data* tokens = nullptr;  // [B][224]
    DecWinState* win = nullptr;        // [B]
    int32_t* done_count = nullptr;
    int32_t* pos_dev = nullptr;        // device-resident step counter read by the graph-replayed decode step
    // beam search: ancestry tables [rows][448] (row-major, double-buffered), per-row limits / history summaries / candidates / parents
    int32_t* beam_anc[2] = {nullptr, nullptr};
    int32_t* beam_anc_cur = nullptr;
    int32_t* beam_limit = nullptr;
    BeamRow* beam_rows = nullptr;
    BeamCand* beam_cands = nullptr;
    int32_t* beam_parent = nullptr;
    float* beam_nosp = nullptr;
    int32_t* beam_rowwin = nullptr;    // [128] row -> window whose cross cache it reads
    int beam_K = 1;                    // rows per window of the current beam / fallback pass (rows w*K .. w*K+K-1 decode window w of the pass)
    cudaGraphExec_t step_graph = nullptr;  // one greedy iteration (sample, advance, step), see decoder_decode_graph
    int graph_B = 0, graph_nodes = 0;
    SampleParams graph_sp{};
    int32_t* ahead_map = nullptr;      // [L*H] -> alignment-head index or -1
    int n_aheads = 0;
    // alignment-head capture of the DTW pass: window b's w[H_a][T_b][A_b] starts at aw + aw_off[b]
    float* aw = nullptr;
    size_t aw_cap = 0;
    int64_t* aw_off = nullptr;
    int32_t* aw_T = nullptr;
    int32_t* aw_A = nullptr;
    unsigned long long* cross_stats = nullptr;  // [2] dec_cross_attn_kernel: launches, live (launch, window) pairs since the last reset
    int cap_W = 0;                     // windows the cross-side buffers (enc_bf16, ckv) hold; cap_B counts ROWS
    int reserve(const wdr_context* ctx, int windows, int rows = 0);  // rows = 0: one row per window
    void release();
};

// Workspace of the batched DTW pass (decoder_dtw_pass): all teacher-forced tokens of all windows as packed rows.
struct DtwPassWorkspace {
    int64_t cap_rows = 0;
    int d = 0;
    float* x = nullptr;            // [M][d] residual stream
    __nv_bfloat16* h = nullptr;    // [2][M][d]  (hi, lo)
    __nv_bfloat16* att = nullptr;  // [2][M][d]
    __nv_bfloat16* ff = nullptr;   // [2][M][4d]
    float* part = nullptr;         // [M][4d] GEMM output (fp32)
    int32_t* row_b = nullptr;      // [M] window of each packed row
    int32_t* row_pos = nullptr;    // [M] token position of each packed row
    int32_t* row_off = nullptr;    // [kDecMaxWindows] first packed row of each window
    int reserve(int64_t rows, int d_model);
    void release();
};

// cross-KV projection of all decoder layers from ws.enc_bf16 (B windows)
int decoder_cross_kv(const wdr_context* ctx, DecoderWorkspace& ws, int B, cudaStream_t st, Profiler* prof);
// one decoder step at position pos for all B windows (token = ws.seq[b][pos]).  want_logits: final LN + logits GEMM.
// mode: DEC_MODE_DECODE = greedy decode (attention skips windows whose DecWinState is completed / failed);
//       DEC_MODE_FORCED = teacher-forced, every window runs;
//       DEC_MODE_DTW    = teacher-forced DTW pass: alignment-head cross-attention rows go to ws.aw, windows stop at their own
//                         length ws.aw_T[b], and without logits only the layers up to the last alignment head run.
//       DEC_MODE_BEAM   = beam search / temperature fallback: row b decodes window ws.beam_rowwin[b] (shared cross cache), the self cache of a
//                         row is read through the ancestry table ws.beam_anc_cur, rows with ws.beam_limit[b] <= pos are skipped.
enum { DEC_MODE_DECODE = 0, DEC_MODE_FORCED = 1, DEC_MODE_DTW = 2, DEC_MODE_BEAM = 3 };
// temperature > 0: logits are divided by it first (whisper_process_logits); probs_out (optional, [R][ws.ldv]) receives the processed
// log-probabilities of every active row (what whisper_sample_token draws from at temperature > 0) and candidate slot k_top the raw
// tid / pt / ptsum of the distribution
int decoder_topk(const wdr_context* ctx, DecoderWorkspace& ws, int R, const SampleParams& sp, int k_top, float temperature, float* probs_out,
                 cudaStream_t st, Profiler* prof);
int decoder_beam_reorder(DecoderWorkspace& ws, int R, int pos_last, cudaStream_t st);
int decoder_step(const wdr_context* ctx, DecoderWorkspace& ws, int B, int pos, bool want_logits, int mode, cudaStream_t st, Profiler* prof,
                 bool pos_on_device = false, bool pdl = false);
// The DTW pass in one shot: the teacher-forced sequences ws.seq[b][0 .. T_b) (T_b = ws.aw_T[b], host copy in T_host; 0 = window
// not in the pass) of all B windows run through the decoder layers up to the last alignment head as ONE batch of sum(T_b) packed
// rows — every linear layer is a full-size tcgen05 GEMM (weights read once, not once per token position) and the cross-attention
// reads each window's K_c/V_c once per 16 queries instead of once per token.  Alignment-head probabilities land in ws.aw exactly as
// in DEC_MODE_DTW of decoder_step (same layout, same fp32 softmax, same (hi, lo) bf16 activations).
int decoder_dtw_pass(const wdr_context* ctx, DecoderWorkspace& ws, DtwPassWorkspace& pw, int B, const int32_t* T_host, cudaStream_t st, Profiler* prof);
// whisper_process_logits + greedy whisper_sample_token + decoder bookkeeping on ws.logits; appends to ws.tokens / ws.seq[pos+1]
int decoder_sample(const wdr_context* ctx, DecoderWorkspace& ws, int B, int pos, const SampleParams& sp, cudaStream_t st, Profiler* prof,
                   bool pos_on_device = false, bool pdl = false);
int decoder_decode_graph(const wdr_context* ctx, DecoderWorkspace& ws, int B, const SampleParams& sp, cudaStream_t st, cudaGraphExec_t* out);

}  // namespace wdr
