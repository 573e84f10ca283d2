// decoder.cu — the KV-cached Whisper text decoder, batched over 30 s windows (one token per window per step).
//
// Replaces whisper.cpp `whisper_build_graph_cross` / `whisper_decode_internal` / `whisper_process_logits` /
// `whisper_sample_token`, which `state.full` runs per window (reference src/transcribe.rs:389,
// configured by setup_params src/transcribe.rs:20-87).
//
// A step advances all B <= 128 windows by one token, so every linear layer is a weight-streaming GEMM with a single
// 128-row M tile: the tcgen05 GEMM runs with BN = 64 tiles and split-K so that >= ~148 CTAs stream the weights, writing
// fp32 partial sums that the consumer kernel reduces in a fixed order (deterministic):
//   dec_ln        x += bias + sum(partials); h = bf16(LayerNorm(x))                      one CTA per window
//   dec_self_attn q|k|v = sum(partials) + bias; append k,v (bf16) to the self cache; softmax(q K^T) V over <= 448 positions
//   dec_cross_attn q = sum(partials) + bias; streams K_c / V_c (2 x 1500 x 64 bf16 per (window, head)) with 128-bit loads —
//                 the HBM-bound kernel of the decoder; optionally stores the alignment heads' probabilities (DTW pass)
//   dec_bias_gelu ff = bf16(gelu(sum(partials) + bias))
//   dec_sample    logit rules, log-softmax, timestamp-mass rule, greedy argmax, token statistics, per-window bookkeeping
// Precision: weights and the cross-KV cache are bf16 (whisper.cpp: f16).  Activations that feed a GEMM (LayerNorm, attention and
// GELU outputs) are stored as a (hi, lo) bf16 pair = 16 mantissa bits and both halves multiply the same weight tile (dual-A
// GEMM: weights are read once, so the weight-streaming cost is unchanged); the self-KV cache, the residual stream, softmax and
// LayerNorm statistics are fp32.  A single-bf16 activation path decorrelates from an fp32 decoder within a few layers (every
// rounding flip is a 2^-8 kick that the next rounding amplifies), which makes greedy argmax flip on ~3 % of the tokens of a
// random-weight model; at 16 bits the logits agree with the fp32 restatement to ~1e-5 and token sequences match.
#include <math.h>
#include <stdlib.h>
#include <algorithm>
#include "common.cuh"
#include "decoder.cuh"
#include "gemm.cuh"

namespace wdr {

constexpr int kT = WDR_AUDIO_CTX;

// ---------------------------------------------------------------------------------------------------
// block-wide reductions (blockDim.x multiple of 32, <= 1024)
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum(float v, float* red) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    float t = (l < nw) ? red[l] : 0.0f;
    t = warp_sum(t);
    return t;
}
__device__ __forceinline__ float block_max(float v, float* red) {
    v = warp_max(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    float t = (l < nw) ? red[l] : -INFINITY;
    t = warp_max(t);
    return t;
}
__device__ __forceinline__ unsigned long long block_max_u64(unsigned long long v, unsigned long long* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long u = __shfl_xor_sync(0xffffffffu, v, o);
        v = u > v ? u : v;
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    unsigned long long t = (l < nw) ? red[l] : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long u = __shfl_xor_sync(0xffffffffu, t, o);
        t = u > t ? u : t;
    }
    return t;
}

__device__ __forceinline__ float gelu_tanh_exact(float x) {
    return 0.5f * x * (1.0f + tanhf(0.79788456080286535587989211986876f * x * (1.0f + 0.044715f * x * x)));
}

// activation v as a (hi, lo) bf16 pair: hi at dst[idx], lo at dst[idx + lo_off]
__device__ __forceinline__ void store_split(__nv_bfloat16* __restrict__ dst, int64_t lo_off, int64_t idx, float v) {
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    dst[idx] = hi;
    dst[idx + lo_off] = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// sum of the split-K partials of element (row, col), in split order
__device__ __forceinline__ float part_sum(const float* __restrict__ part, int n_splits, int64_t split_stride, int64_t idx) {
    const float* pp = part + idx;
    float a = pp[0];
    int s = 1;
    for (; s + 8 <= n_splits; s += 8) {  // eight loads in flight (one L2 round trip for the usual 7-9 splits), summed in split order
        float t[8];
#pragma unroll
        for (int k = 0; k < 8; k++) t[k] = pp[(int64_t)(s + k) * split_stride];
#pragma unroll
        for (int k = 0; k < 8; k++) a += t[k];
    }
    if (s < n_splits) {  // the tail: still all loads first, then the adds in split order
        float t[7];
#pragma unroll
        for (int k = 0; k < 7; k++) t[k] = (s + k < n_splits) ? pp[(int64_t)(s + k) * split_stride] : 0.0f;
#pragma unroll
        for (int k = 0; k < 7; k++) if (s + k < n_splits) a += t[k];
    }
    return a;
}

// ---------------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------------
// Position arguments come as (pos_ptr, pos): a non-null pos_ptr (the workspace's device-resident step counter) wins, so that a
// captured decode step replays as a CUDA graph without patching kernel arguments.
__device__ __forceinline__ int load_pos(const int32_t* __restrict__ pos_ptr, int pos) { return pos_ptr ? *pos_ptr : pos; }

__global__ void dec_advance_kernel(int32_t* pos_ptr) {
    pdl_launch_dependents();
    pdl_wait();
    *pos_ptr += 1;
}

__global__ void dec_embed_kernel(const int32_t* __restrict__ seq, const int32_t* __restrict__ pos_ptr, int pos, const __nv_bfloat16* __restrict__ tok_emb,
                                 const float* __restrict__ pos_emb, int d, int n_vocab, float* __restrict__ x) {
    pdl_launch_dependents();
    pdl_wait();
    pos = load_pos(pos_ptr, pos);
    const int b = blockIdx.x;
    int tok = seq[b * kDecSeqCap + pos];
    if (tok < 0 || tok >= n_vocab) tok = 0;
    for (int i = threadIdx.x; i < d; i += blockDim.x)
        x[(int64_t)b * d + i] = __bfloat162float(tok_emb[(int64_t)tok * d + i]) + pos_emb[(int64_t)pos * d + i];
}

// x (+= bias + partials) -> h = (hi, lo) bf16 of LN(x).  One CTA per window row, one float4 column group per thread
// (blockDim.x = d / 4 <= 320): the n_splits partial loads of a thread are independent and issue back to back.
__global__ void __launch_bounds__(320)
dec_ln_kernel(float* __restrict__ x, const float* __restrict__ part, int n_splits, int64_t split_stride, int ldp,
              const float* __restrict__ bias, const float* __restrict__ g, const float* __restrict__ bta, __nv_bfloat16* __restrict__ h, int64_t lo_off,
              int d) {
    __shared__ float red[32];
    pdl_launch_dependents();
    const int row = blockIdx.x, i = threadIdx.x * 4;
    const float4 gg = *reinterpret_cast<const float4*>(g + i), bb = *reinterpret_cast<const float4*>(bta + i);  // weights: no dependency
    pdl_wait();
    float4 a = *reinterpret_cast<const float4*>(x + (int64_t)row * d + i);
    if (part) {
        const float* pp = part + (int64_t)row * ldp + i;
        float4 acc = *reinterpret_cast<const float4*>(pp);
        int s = 1;
        for (; s + 8 <= n_splits; s += 8) {  // eight loads in flight, summed in split order
            float4 t[8];
#pragma unroll
            for (int k = 0; k < 8; k++) t[k] = *reinterpret_cast<const float4*>(pp + (int64_t)(s + k) * split_stride);
#pragma unroll
            for (int k = 0; k < 8; k++) { acc.x += t[k].x; acc.y += t[k].y; acc.z += t[k].z; acc.w += t[k].w; }
        }
        if (s < n_splits) {  // the tail (<= 7 splits): all loads first — one L2 round trip —, then the adds in split order
            float4 t[7];
#pragma unroll
            for (int k = 0; k < 7; k++)
                t[k] = (s + k < n_splits) ? *reinterpret_cast<const float4*>(pp + (int64_t)(s + k) * split_stride) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int k = 0; k < 7; k++)
                if (s + k < n_splits) { acc.x += t[k].x; acc.y += t[k].y; acc.z += t[k].z; acc.w += t[k].w; }
        }
        const float4 bs = *reinterpret_cast<const float4*>(bias + i);
        a.x += acc.x + bs.x; a.y += acc.y + bs.y; a.z += acc.z + bs.z; a.w += acc.w + bs.w;
        *reinterpret_cast<float4*>(x + (int64_t)row * d + i) = a;
    }
    const float mean = block_sum(a.x + a.y + a.z + a.w, red) / (float)d;
    const float c0 = a.x - mean, c1 = a.y - mean, c2 = a.z - mean, c3 = a.w - mean;
    const float var = block_sum(c0 * c0 + c1 * c1 + c2 * c2 + c3 * c3, red) / (float)d;
    const float rstd = 1.0f / sqrtf(var + 1e-5f);
    const int64_t o = (int64_t)row * d + i;
    store_split(h, lo_off, o, c0 * rstd * gg.x + bb.x);
    store_split(h, lo_off, o + 1, c1 * rstd * gg.y + bb.y);
    store_split(h, lo_off, o + 2, c2 * rstd * gg.z + bb.z);
    store_split(h, lo_off, o + 3, c3 * rstd * gg.w + bb.w);
}

// self-attention at position pos: one WARP per (head, window), one warp per CTA — no block-wide barrier anywhere.
// part: [S][B][3d] partials of the fused QKV GEMM.
//
// Self cache: f16 (whisper.cpp keeps kv_self in GGML_TYPE_F16 as well), head-major, [window][head] blocks of 448 x 64.  V is
// row-major [t][64] (128 B rows).  K is stored TRANSPOSED IN BLOCKS OF 32 POSITIONS — [t / 32][c / 8][t % 32][c % 8] — so that
// "lane = key position" reads of a block are conflict-free 16-byte pieces, in global and in shared memory alike.
//
// The kernel is bound by its chain of dependent memory round trips, not by bytes (ncu, large-v3, 120 windows, in-graph): with an fp32
// cache read straight into registers it took 15 us at pos 3 and 34-40 us at pos 110 (one round trip per 32 keys, then one per 8-16
// value rows; 212 ms of a 1623 ms decode), and an f16 cache alone changed nothing — under a register budget that keeps all 2400
// (window, head) warps of a 120-window batch resident, ptxas serialises the "batched" loads.  So the history does not go through
// registers at all: the warp streams it as 4 KB chunks (32 positions of K, then 32 rows of V) with cp.async into a two-buffer ring
// in shared memory — requested first thing, before the q / k / v partial sums of the new position are even fetched — and computes
// each chunk from shared memory.  The new position's k / v never make the global round trip: they are patched into the chunk that
// holds them.  Scores: lane = key position (one fmaf chain in column order); softmax by warp shuffles; P V: lane = two adjacent
// columns, one fmaf chain in position order.
// Four heads (warps) per CTA: with one-warp CTAs the 2400 CTA launches of a 120-window large-v3 batch (16 per SM, one after the other)
// were themselves a large part of the kernel's ~16 us floor (ncu, L2-warm, pos 110: 16.6 us).
constexpr int kSelfMaxHeadsPerCta = 4;
__device__ __forceinline__ int64_t self_k_off(int t, int c) { return (int64_t)(t >> 5) * 2048 + (c >> 3) * 256 + (t & 31) * 8 + (c & 7); }
__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// ANC = beam search / fallback rows (history read through the ancestry table); the greedy instantiation carries none of that code.
template <bool ANC>
__global__ void __launch_bounds__(kSelfMaxHeadsPerCta * 32)
dec_self_attn_kernel(const float* __restrict__ part, int n_splits, int64_t split_stride, const float* __restrict__ b_qkv,
                     __half* __restrict__ sk, __half* __restrict__ sv, const int32_t* __restrict__ pos_ptr, int pos, int d, __nv_bfloat16* __restrict__ att, int64_t lo_off,
                     const DecWinState* __restrict__ win /* decode: skip finished windows */, const int32_t* __restrict__ t_limit /* forced pass: window length */,
                     const int32_t* __restrict__ anc /* beam search: [rows][anc_ld = 448] row that holds position t of this row's history; null = own row */,
                     int anc_ld) {
    __shared__ __align__(16) unsigned char ring_[kSelfMaxHeadsPerCta][2][4096];
    __shared__ __align__(16) float q_[kSelfMaxHeadsPerCta][64];
    __shared__ float p_[kSelfMaxHeadsPerCta][kDecSeqCap];
    __shared__ int32_t as__[ANC ? kSelfMaxHeadsPerCta : 1][ANC ? kDecSeqCap : 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int hh = blockIdx.x * (blockDim.x >> 5) + warp, b = blockIdx.y;
    unsigned char (*ring)[4096] = ring_[warp];
    float* q = q_[warp];
    float* p = p_[warp];
    int32_t* as_ = as__[ANC ? warp : 0];
    pdl_launch_dependents();
    pdl_wait();
    pos = load_pos(pos_ptr, pos);
    if (win && (win[b].completed | win[b].failed)) return;
    if (t_limit && pos >= t_limit[b]) return;
    const int n_heads = d >> 6;
    __half* K = sk + ((int64_t)b * n_heads + hh) * kDecSeqCap * 64;
    __half* V = sv + ((int64_t)b * n_heads + hh) * kDecSeqCap * 64;
    // beam search: position t < pos of this row's history lives in the cache of row anc[t][b] (the beam it descended from); the
    // table is row-major [row][448], this row's ancestry is copied into shared memory once (coalesced)
    const int64_t row_step = (int64_t)n_heads * kDecSeqCap * 64;
    if (ANC) {
        for (int t = lane; t < kDecSeqCap; t += 32) as_[t] = t < pos ? anc[(int64_t)b * anc_ld + t] : b;
        __syncwarp();
    }
#define WDR_ANC_ROW(t) (ANC ? (int64_t)(as_[(t)] - b) * row_step : (int64_t)0)
    // chunk i of the stream: i < nc -> K block i, else V rows 32 (i - nc) .. + 31 (positions beyond pos are copied too: cache memory
    // of this (window, head), never used)
    const int nc = (pos >> 5) + 1;
    auto issue = [&](int i) {
        if (i < 2 * nc) {
            unsigned char* dst = ring[i & 1];
            if (i < nc) {
                const int t = 32 * i + lane;
                const __half* src = K + WDR_ANC_ROW(t) + (int64_t)i * 2048 + lane * 8;
#pragma unroll
                for (int c8 = 0; c8 < 8; c8++) cp_async_16(dst + c8 * 512 + lane * 16, src + c8 * 256);
            } else {
                const int j = i - nc;
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const int r = k * 4 + (lane >> 3), t = 32 * j + r;
                    cp_async_16(dst + r * 128 + (lane & 7) * 16, V + WDR_ANC_ROW(t) + (int64_t)t * 64 + (lane & 7) * 8);
                }
            }
        }
        cp_async_commit();  // (an empty group when the stream has ended: the wait below always allows exactly one pending group)
    };
    issue(0);
    issue(1);
    // q / k / v of the new position: all partial loads of a split batch are issued together, summed in split order
    float qkv[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    {
        const float* pp = part + (int64_t)b * 3 * d + hh * 64 + lane;
        for (int s0 = 0; s0 < n_splits; s0 += 4) {
            float t[4][6];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const float* ps = pp + (int64_t)min(s0 + k, n_splits - 1) * split_stride;
#pragma unroll
                for (int m = 0; m < 6; m++) t[k][m] = ps[(m >> 1) * d + (m & 1) * 32];
            }
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (s0 + k < n_splits) {
#pragma unroll
                    for (int m = 0; m < 6; m++) qkv[m] += t[k][m];
                }
        }
    }
    __half kn[2], vn[2];
#pragma unroll
    for (int m = 0; m < 2; m++) {
        const int e = lane + 32 * m;
        q[e] = qkv[m] + b_qkv[hh * 64 + e];
        kn[m] = __float2half_rn(qkv[2 + m] + b_qkv[d + hh * 64 + e]);
        vn[m] = __float2half_rn(qkv[4 + m] + b_qkv[2 * d + hh * 64 + e]);
        K[self_k_off(pos, e)] = kn[m];
        V[(int64_t)pos * 64 + e] = vn[m];
    }
    __syncwarp();
    const float4* qv = reinterpret_cast<const float4*>(q);
    float mx = -INFINITY, inv = 0.0f, a0 = 0.0f, a1 = 0.0f;
    for (int i = 0; i < 2 * nc; i++) {
        cp_async_wait<1>();
        __syncwarp();
        unsigned char* buf = ring[i & 1];
        if (i == nc - 1 || i == 2 * nc - 1) {  // the chunk that holds the new position: patch k / v in (the copy brought stale memory)
#pragma unroll
            for (int m = 0; m < 2; m++) {
                const int e = lane + 32 * m;
                if (i < nc) *reinterpret_cast<__half*>(buf + (e >> 3) * 512 + (pos & 31) * 16 + (e & 7) * 2) = kn[m];
                else *reinterpret_cast<__half*>(buf + (pos & 31) * 128 + e * 2) = vn[m];
            }
            __syncwarp();
        }
        if (i < nc) {
            // ---- scores of 32 positions: lane = position ----
            const int t = 32 * i + lane;
            float a = 0.0f;
#pragma unroll
            for (int c8 = 0; c8 < 8; c8++) {
                const uint4 u = *reinterpret_cast<const uint4*>(buf + c8 * 512 + lane * 16);
                const float4 qa = qv[2 * c8], qb = qv[2 * c8 + 1];
                const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&u.x)), f1 = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
                const float2 f2 = __half22float2(*reinterpret_cast<const __half2*>(&u.z)), f3 = __half22float2(*reinterpret_cast<const __half2*>(&u.w));
                a = fmaf(qa.x, f0.x, a); a = fmaf(qa.y, f0.y, a); a = fmaf(qa.z, f1.x, a); a = fmaf(qa.w, f1.y, a);
                a = fmaf(qb.x, f2.x, a); a = fmaf(qb.y, f2.y, a); a = fmaf(qb.z, f3.x, a); a = fmaf(qb.w, f3.y, a);
            }
            a *= 0.125f;
            if (t <= pos) { p[t] = a; mx = fmaxf(mx, a); }
            if (i == nc - 1) {  // all scores are in: softmax
                mx = warp_max(mx);
                __syncwarp();
                float sum = 0.0f;
                for (int tt = lane; tt <= pos; tt += 32) {
                    const float e = expf(p[tt] - mx);
                    p[tt] = e;
                    sum += e;
                }
                sum = warp_sum(sum);
                inv = 1.0f / sum;
            }
        } else {
            // ---- P V over 32 rows: lane = columns 2 lane, 2 lane + 1 ----
            const int t0 = 32 * (i - nc), n = min(32, pos + 1 - t0);
            for (int r = 0; r < n; r++) {
                const float pt = p[t0 + r] * inv;
                const float2 f = __half22float2(*reinterpret_cast<const __half2*>(buf + r * 128 + lane * 4));
                a0 = fmaf(pt, f.x, a0);
                a1 = fmaf(pt, f.y, a1);
            }
        }
        __syncwarp();  // every lane is done with this buffer before the next chunk lands in it
        issue(i + 2);
    }
    {   // (hi, lo) of the two adjacent columns as packed stores
        const __nv_bfloat162 hi = __floats2bfloat162_rn(a0, a1);
        const float2 hf = __bfloat1622float2(hi);
        const __nv_bfloat162 lo = __floats2bfloat162_rn(a0 - hf.x, a1 - hf.y);
        __nv_bfloat16* o = att + (int64_t)b * d + hh * 64 + 2 * lane;
        *reinterpret_cast<__nv_bfloat162*>(o) = hi;
        *reinterpret_cast<__nv_bfloat162*>(o + lo_off) = lo;
    }
#undef WDR_ANC_ROW
}

// cross-attention of one (head, window): q from the cross-query GEMM partials; K_c/V_c rows are 64 bf16 (128 B) at row
// stride 2d.  8 lanes x 16 B cover one row; a warp covers 4 rows per load, the CTA (8 warps) 32 rows.
template <int MINB, bool ROWMAP>
__global__ void __launch_bounds__(256, MINB)
dec_cross_attn_kernel(const float* __restrict__ part, int n_splits, int64_t split_stride, const float* __restrict__ b_q,
                      const __nv_bfloat16* __restrict__ ckv, int d, __nv_bfloat16* __restrict__ att, int64_t lo_off,
                      const int32_t* __restrict__ ahead_map /* this layer's [H] -> alignment-head index or -1; null = no capture */,
                      float* __restrict__ aw, const int64_t* __restrict__ aw_off, const int32_t* __restrict__ aw_T,
                      const int32_t* __restrict__ aw_A, const int32_t* __restrict__ pos_ptr, int pos, const DecWinState* __restrict__ win,
                      const int32_t* __restrict__ t_limit, const int32_t* __restrict__ row_window /* beam search / fallback: row b reads the cross cache of window row_window[b]; null = b */,
                      unsigned long long* __restrict__ stats /* optional [2]: launches, (launch, window) pairs that really streamed their K_c / V_c */) {
    __shared__ float q[64];
    __shared__ float p[kT + 4];
    __shared__ float red[32];
    __shared__ float accs[8][64];
    const int hh = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5, g = lane & 7, r = lane >> 3;
    pdl_launch_dependents();
    pdl_wait();
    pos = load_pos(pos_ptr, pos);
    if (stats && tid == 0 && hh == 0 && b == 0) atomicAdd(&stats[0], 1ull);
    if (win && (win[b].completed | win[b].failed)) return;  // finished windows stop streaming their K/V
    if (t_limit && pos >= t_limit[b]) return;
    if (stats && tid == 0 && hh == 0) atomicAdd(&stats[1], 1ull);  // bench.py: algorithmic bytes follow the LIVE windows of a launch
    if (tid < 64) q[tid] = part_sum(part, n_splits, split_stride, (int64_t)b * d + hh * 64 + tid) + b_q[hh * 64 + tid];
    __syncthreads();
    float q8[8];
#pragma unroll
    for (int j = 0; j < 8; j++) q8[j] = q[g * 8 + j];
    // head-major cross cache: [(window, head)][K | V][1500][64] — both blocks of a CTA are contiguous 192 KB streams
    const __nv_bfloat16* Kb = ckv + ((int64_t)(ROWMAP ? row_window[b] : b) * gridDim.x + hh) * 2 * kT * 64 + g * 8;
    const __nv_bfloat16* Vb = Kb + kT * 64;
    const int64_t rs = 64;
    // ---- scores ----
    float mx = -INFINITY;
    for (int t0 = warp * 4 + r; t0 < kT; t0 += 128) {
        uint4 u[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int t = t0 + k * 32;
            u[k] = (t < kT) ? __ldg(reinterpret_cast<const uint4*>(Kb + (int64_t)t * rs)) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int t = t0 + k * 32;
            const __nv_bfloat162* e = reinterpret_cast<const __nv_bfloat162*>(&u[k]);
            float a = 0.0f;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float2 f = __bfloat1622float2(e[j]);
                a = fmaf(q8[2 * j], f.x, a);
                a = fmaf(q8[2 * j + 1], f.y, a);
            }
            a += __shfl_xor_sync(0xffffffffu, a, 1);
            a += __shfl_xor_sync(0xffffffffu, a, 2);
            a += __shfl_xor_sync(0xffffffffu, a, 4);
            a *= 0.125f;
            if (t < kT) {
                if (g == 0) p[t] = a;
                mx = fmaxf(mx, a);
            }
        }
    }
    mx = block_max(mx, red);
    float sum = 0.0f;
    for (int t = tid; t < kT; t += 256) {
        const float e = expf(p[t] - mx);
        p[t] = e;
        sum += e;
    }
    sum = block_sum(sum, red);
    const float inv = 1.0f / sum;
    for (int t = tid; t < kT; t += 256) p[t] *= inv;
    __syncthreads();
    const int ahead = ahead_map ? ahead_map[hh] : -1;
    if (ahead >= 0) {
        const int T_b = aw_T[b], A_b = aw_A[b];
        if (pos < T_b) {
            float* dst = aw + aw_off[b] + ((int64_t)ahead * T_b + pos) * A_b;
            for (int t = tid; t < A_b; t += 256) dst[t] = p[t];
        }
    }
    // ---- P V ----
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; j++) acc[j] = 0.0f;
    for (int t0 = warp * 4 + r; t0 < kT; t0 += 128) {
        uint4 u[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int t = t0 + k * 32;
            u[k] = (t < kT) ? __ldg(reinterpret_cast<const uint4*>(Vb + (int64_t)t * rs)) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int t = t0 + k * 32;
            const float pt = (t < kT) ? p[t] : 0.0f;
            const __nv_bfloat162* e = reinterpret_cast<const __nv_bfloat162*>(&u[k]);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float2 f = __bfloat1622float2(e[j]);
                acc[2 * j] = fmaf(pt, f.x, acc[2 * j]);
                acc[2 * j + 1] = fmaf(pt, f.y, acc[2 * j + 1]);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 8; j++) {
        acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 8);
        acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 16);
    }
    if (r == 0) {
#pragma unroll
        for (int j = 0; j < 8; j++) accs[warp][g * 8 + j] = acc[j];
    }
    __syncthreads();
    if (tid < 64) {
        float a = 0.0f;
#pragma unroll
        for (int w = 0; w < 8; w++) a += accs[w][tid];
        store_split(att, lo_off, (int64_t)b * d + hh * 64 + tid, a);
    }
}

// Beam search / best_of decoding: the K rows (decoders) of a window share its K_c / V_c.  One CTA per (head, window) reads the two
// 192 KB blocks ONCE for all its rows (the single-query kernel re-reads them per row and is L2-bound at beam 5: 113 us per launch
// against 30 us of HBM time).  Same arithmetic as the batched DTW-pass kernel: thread = key row, the queries (pre-scaled by 1/8,
// interleaved in pairs) broadcast from shared memory into packed FFMA2s, one fmaf chain per (row, query) in column order; softmax
// by one warp per row; P V with warp = key slice, lane = column pair.  NQP = query pairs (rows padded to 2 NQP with zero queries).
// Rows b = w * K + i; rows with t_limit[b] <= pos are dead: nothing is stored for them; a window with no live row returns at once.
template <int NQP>
__global__ void __launch_bounds__(256, 4)  // four CTAs per SM: the 500 CTAs of a 25-window beam-5 batch are one wave
dec_cross_attn_rows_kernel(const float* __restrict__ part, int n_splits, int64_t split_stride, const float* __restrict__ b_q,
                           const __nv_bfloat16* __restrict__ ckv, int d, __nv_bfloat16* __restrict__ att, int64_t lo_off, int pos,
                           const int32_t* __restrict__ t_limit, const int32_t* __restrict__ row_window, int K) {
    extern __shared__ __align__(16) float rows_smem[];
    constexpr int NQ = 2 * NQP;
    constexpr int kPS = kT + 4;
    float* qs = rows_smem;        // [NQP][64 columns][2 queries]
    float* p = qs + NQP * 128;    // [NQ][kT + 4]; reused as the P V reduction buffer [8][NQ][64]
    const int hh = blockIdx.x, w = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b0 = w * K;
    bool any = false;
    for (int i = 0; i < K; i++) any |= pos < t_limit[b0 + i];
    if (!any) return;
    for (int e = tid; e < NQP * 128; e += 256) {
        const int qp = e >> 7, c = (e & 127) >> 1, qi = 2 * qp + (e & 1);
        qs[e] = qi < K ? (part_sum(part, n_splits, split_stride, (int64_t)(b0 + qi) * d + hh * 64 + c) + b_q[hh * 64 + c]) * 0.125f : 0.0f;
    }
    for (int e = tid; e < NQ * 4; e += 256) p[(e >> 2) * kPS + kT + (e & 3)] = 0.0f;
    __syncthreads();
    const __nv_bfloat16* Kb = ckv + ((int64_t)row_window[b0] * gridDim.x + hh) * 2 * kT * 64;
    const __nv_bfloat16* Vb = Kb + kT * 64;
    // ---- scores: two key rows per thread, columns outer, query pairs inner ----
    for (int t = tid; t < kT; t += 512) {
        const int t1 = t + 256;
        const bool has1 = t1 < kT;
        uint4 ka[8], kb[8];
        {
            const uint4* kra = reinterpret_cast<const uint4*>(Kb + (int64_t)t * 64);
            const uint4* krb = reinterpret_cast<const uint4*>(Kb + (int64_t)(has1 ? t1 : t) * 64);
#pragma unroll
            for (int c8 = 0; c8 < 8; c8++) { ka[c8] = __ldg(kra + c8); kb[c8] = __ldg(krb + c8); }
        }
        float2 a0[NQP], a1[NQP];
#pragma unroll
        for (int qp = 0; qp < NQP; qp++) { a0[qp] = make_float2(0.0f, 0.0f); a1[qp] = make_float2(0.0f, 0.0f); }
#pragma unroll
        for (int c8 = 0; c8 < 8; c8++) {
            const __nv_bfloat162* ea = reinterpret_cast<const __nv_bfloat162*>(&ka[c8]);
            const __nv_bfloat162* eb = reinterpret_cast<const __nv_bfloat162*>(&kb[c8]);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float2 fa = __bfloat1622float2(ea[j]), fb = __bfloat1622float2(eb[j]);
#pragma unroll
                for (int qp = 0; qp < NQP; qp++) {
                    const float4 f = *reinterpret_cast<const float4*>(qs + qp * 128 + (c8 * 4 + j) * 4);
                    a0[qp] = __ffma2_rn(make_float2(f.x, f.y), make_float2(fa.x, fa.x), a0[qp]);
                    a0[qp] = __ffma2_rn(make_float2(f.z, f.w), make_float2(fa.y, fa.y), a0[qp]);
                    a1[qp] = __ffma2_rn(make_float2(f.x, f.y), make_float2(fb.x, fb.x), a1[qp]);
                    a1[qp] = __ffma2_rn(make_float2(f.z, f.w), make_float2(fb.y, fb.y), a1[qp]);
                }
            }
        }
#pragma unroll
        for (int qp = 0; qp < NQP; qp++) {
            p[(2 * qp) * kPS + t] = a0[qp].x;
            p[(2 * qp + 1) * kPS + t] = a0[qp].y;
            if (has1) {
                p[(2 * qp) * kPS + t1] = a1[qp].x;
                p[(2 * qp + 1) * kPS + t1] = a1[qp].y;
            }
        }
    }
    __syncthreads();
    // ---- softmax: warp i = row i ----
    if (warp < K) {
        float* pr = p + warp * kPS;
        float mx = -INFINITY;
        for (int t = lane; t < kT; t += 32) mx = fmaxf(mx, pr[t]);
        mx = warp_max(mx);
        float sum = 0.0f;
        for (int t = lane; t < kT; t += 32) {
            const float e = expf(pr[t] - mx);
            pr[t] = e;
            sum += e;
        }
        sum = warp_sum(sum);
        const float inv = 1.0f / sum;
        for (int t = lane; t < kT; t += 32) pr[t] *= inv;
    }
    __syncthreads();
    // ---- P V: warp = key slice (groups of 4 keys), lane = column pair ----
    float2 acc[NQ];
#pragma unroll
    for (int qi = 0; qi < NQ; qi++) acc[qi] = make_float2(0.0f, 0.0f);
    // eight keys per trip (eight independent 128 B row loads per warp in flight); kT = 1500 = 8 * 187 + 4: the last trip is half
    for (int t8 = warp * 8; t8 < kT; t8 += 64) {
        const bool second = t8 + 4 < kT;  // warp-uniform
        float2 v[8];
#pragma unroll
        for (int j = 0; j < 8; j++)
            v[j] = (j < 4 || second) ? __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(Vb + (int64_t)(t8 + j) * 64 + 2 * lane)) : make_float2(0.0f, 0.0f);
#pragma unroll
        for (int g4 = 0; g4 < 2; g4++) {
            if (g4 == 1 && !second) break;
#pragma unroll
            for (int qi = 0; qi < NQ; qi++) {
                const float4 pv = *reinterpret_cast<const float4*>(p + qi * kPS + t8 + 4 * g4);
                acc[qi] = __ffma2_rn(make_float2(pv.x, pv.x), v[4 * g4], acc[qi]);
                acc[qi] = __ffma2_rn(make_float2(pv.y, pv.y), v[4 * g4 + 1], acc[qi]);
                acc[qi] = __ffma2_rn(make_float2(pv.z, pv.z), v[4 * g4 + 2], acc[qi]);
                acc[qi] = __ffma2_rn(make_float2(pv.w, pv.w), v[4 * g4 + 3], acc[qi]);
            }
        }
    }
    __syncthreads();  // all warps are done reading p
    float* red = p;   // [8][NQ][64]
#pragma unroll
    for (int qi = 0; qi < NQ; qi++) {
        red[(warp * NQ + qi) * 64 + 2 * lane] = acc[qi].x;
        red[(warp * NQ + qi) * 64 + 2 * lane + 1] = acc[qi].y;
    }
    __syncthreads();
    for (int e = tid; e < K * 64; e += 256) {
        const int qi = e >> 6, c = e & 63;
        if (pos >= t_limit[b0 + qi]) continue;
        float a = 0.0f;
#pragma unroll
        for (int wv = 0; wv < 8; wv++) a += red[(wv * NQ + qi) * 64 + c];
        store_split(att, lo_off, (int64_t)(b0 + qi) * d + hh * 64 + c, a);
    }
}

template <int NQP>
static cudaError_t launch_cross_rows(int H, int nW, int K, cudaStream_t st, const float* part, int n_splits, int64_t split_stride, const float* b_q,
                                     const __nv_bfloat16* ckv, int d, __nv_bfloat16* att, int64_t lo_off, int pos, const int32_t* t_limit,
                                     const int32_t* row_window) {
    const size_t smem = sizeof(float) * ((size_t)NQP * 128 + (size_t)2 * NQP * (kT + 4));
    static DeviceOnce attr_once;
    {
        const cudaError_t e = per_device_once(attr_once, [&] { return cudaFuncSetAttribute(dec_cross_attn_rows_kernel<NQP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); });
        if (e != cudaSuccess) return e;
    }
    dec_cross_attn_rows_kernel<NQP><<<dim3(H, nW), 256, smem, st>>>(part, n_splits, split_stride, b_q, ckv, d, att, lo_off, pos, t_limit, row_window, K);
    return cudaGetLastError();
}

__global__ void dec_bias_gelu_kernel(const float* __restrict__ part, int n_splits, int64_t split_stride, const float* __restrict__ bias,
                                     int n, int64_t total, __nv_bfloat16* __restrict__ out, int64_t lo_off) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int j = (int)(e % n);
        store_split(out, lo_off, e, gelu_tanh_exact(part_sum(part, n_splits, split_stride, e) + bias[j]));
    }
}


// ---------------------------------------------------------------------------------------------------
// batched DTW pass: packed rows (window b, position pos), see decoder.cuh
// ---------------------------------------------------------------------------------------------------
__global__ void dtwp_embed_kernel(const int32_t* __restrict__ seq, const int32_t* __restrict__ row_b, const int32_t* __restrict__ row_pos,
                                  const __nv_bfloat16* __restrict__ tok_emb, const float* __restrict__ pos_emb, int d, int n_vocab,
                                  float* __restrict__ x) {
    const int row = blockIdx.x, b = row_b[row], pos = row_pos[row];
    int tok = seq[b * kDecSeqCap + pos];
    if (tok < 0 || tok >= n_vocab) tok = 0;
    for (int i = threadIdx.x; i < d; i += blockDim.x)
        x[(int64_t)row * d + i] = __bfloat162float(tok_emb[(int64_t)tok * d + i]) + pos_emb[(int64_t)pos * d + i];
}

constexpr int kDtwpMaxT = 256;  // longest teacher-forced sequence (sot, lang, not, <= 220 text tokens, eot)

// causal self-attention of one (head, window) over all its T_b positions; K and V (+bias, f16-rounded) staged in shared memory as fp32
// (row stride 65: conflict-free for "lane = key" and "lane = column" accesses alike); one warp per query.
__global__ void __launch_bounds__(256)
dtwp_self_attn_kernel(const float* __restrict__ qkv /* [M][3d] */, const float* __restrict__ b_qkv, const int32_t* __restrict__ row_off,
                      const int32_t* __restrict__ T, int d, __nv_bfloat16* __restrict__ att, int64_t lo_off, int cap_T /* >= every T_b: sizes the shared arrays */) {
    extern __shared__ float sm[];
    const int hh = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T_b = T[b];
    if (T_b <= 0) return;
    const int64_t r0 = row_off[b];
    float* Ks = sm;                       // [T_b][65]
    float* Vs = Ks + cap_T * 65;          // [T_b][65]
    float* qs = Vs + cap_T * 65;          // [8][64]
    float* ps = qs + 8 * 64;              // [8][cap_T]
    for (int e = tid; e < T_b * 64; e += 256) {
        const int t = e >> 6, c = e & 63;
        const float* src = qkv + (r0 + t) * 3 * (int64_t)d + hh * 64 + c;
        // rounded to f16 as the decode's self cache stores them: the forced pass sees the keys / values the decode saw
        Ks[t * 65 + c] = __half2float(__float2half_rn(src[d] + b_qkv[d + hh * 64 + c]));
        Vs[t * 65 + c] = __half2float(__float2half_rn(src[2 * d] + b_qkv[2 * d + hh * 64 + c]));
    }
    __syncthreads();
    float* q = qs + warp * 64;
    float* p = ps + warp * cap_T;
    for (int i = warp; i < T_b; i += 8) {
        const float* src = qkv + (r0 + i) * 3 * (int64_t)d + hh * 64;
        q[lane] = src[lane] + b_qkv[hh * 64 + lane];
        q[lane + 32] = src[lane + 32] + b_qkv[hh * 64 + lane + 32];
        __syncwarp();
        float mx = -INFINITY;
        for (int j = lane; j <= i; j += 32) {
            const float* kr = Ks + j * 65;
            float a = 0.0f;
#pragma unroll 16
            for (int c = 0; c < 64; c++) a = fmaf(q[c], kr[c], a);
            a *= 0.125f;
            p[j] = a;
            mx = fmaxf(mx, a);
        }
        mx = warp_max(mx);
        float sum = 0.0f;
        for (int j = lane; j <= i; j += 32) {
            const float e = expf(p[j] - mx);
            p[j] = e;
            sum += e;
        }
        sum = warp_sum(sum);
        const float inv = 1.0f / sum;
        __syncwarp();
        float a0 = 0.0f, a1 = 0.0f;
        for (int j = 0; j <= i; j++) {
            const float pj = p[j] * inv;
            a0 = fmaf(pj, Vs[j * 65 + lane], a0);
            a1 = fmaf(pj, Vs[j * 65 + lane + 32], a1);
        }
        const int64_t o = (r0 + i) * (int64_t)d + hh * 64;
        store_split(att, lo_off, o + lane, a0);
        store_split(att, lo_off, o + lane + 32, a1);
        __syncwarp();
    }
}

// cross-attention of 16 queries of one (window, head) against the window's 1500 K_c / V_c rows.  Scores: one key row
// (64 bf16 -> fp32 registers) per thread, the 16 queries broadcast from shared memory; softmax per query by one warp;
// P V with one warp per key slice (probabilities broadcast as float4, 8 FMAs per shared load).  Probabilities of the
// alignment heads are stored to aw (same layout as dec_cross_attn_kernel).
constexpr int kDtwpQB = 16;
constexpr int kDtwpPStride = kT + 4;
__global__ void __launch_bounds__(256)
dtwp_cross_attn_kernel(const float* __restrict__ qpart /* [M][d] */, const float* __restrict__ b_q, const __nv_bfloat16* __restrict__ ckv, int d,
                       const int32_t* __restrict__ row_off, const int32_t* __restrict__ T, __nv_bfloat16* __restrict__ att, int64_t lo_off,
                       const int32_t* __restrict__ ahead_map, float* __restrict__ aw, const int64_t* __restrict__ aw_off,
                       const int32_t* __restrict__ aw_A) {
    extern __shared__ float sm[];
    const int hh = blockIdx.y, b = blockIdx.z, q0 = blockIdx.x * kDtwpQB, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T_b = T[b];
    if (q0 >= T_b) return;
    const int nq = min(kDtwpQB, T_b - q0);
    const int64_t r0 = row_off[b] + q0;
    float* qs = sm;                         // [16][64]
    float* p = qs + kDtwpQB * 64;           // [16][1504]; reused as the P V reduction buffer [8][16][64]
    // queries interleaved in pairs, qs[(qi / 2) * 128 + 2 c + (qi & 1)]: one 64-bit operand holds column c of two queries, so the
    // scores below are packed FFMA2s (fma.rn.f32x2, the key value broadcast) — per query the same fmaf chain in the same order
    for (int e = tid; e < kDtwpQB * 64; e += 256) {
        const int qi = e >> 6, c = e & 63;
        qs[(qi >> 1) * 128 + 2 * c + (qi & 1)] = qi < nq ? (qpart[(r0 + qi) * (int64_t)d + hh * 64 + c] + b_q[hh * 64 + c]) * 0.125f : 0.0f;
    }
    for (int e = tid; e < kDtwpQB * 4; e += 256) p[(e >> 2) * kDtwpPStride + kT + (e & 3)] = 0.0f;
    __syncthreads();
    const __nv_bfloat16* Kb = ckv + ((int64_t)b * gridDim.y + hh) * 2 * kT * 64;
    const __nv_bfloat16* Vb = Kb + kT * 64;
    const int64_t rs = 64;
    // ---- scores: two key rows per thread (kept as packed bf16, 64 registers), columns outer, query pairs inner — every
    // broadcast LDS.128 of the queries (the shared-memory pipe bounds this loop) feeds four FFMA2s instead of two ----
    for (int t = tid; t < kT; t += 512) {
        const int t1 = t + 256;
        const bool has1 = t1 < kT;
        uint4 ka[8], kb[8];
        {
            const uint4* kra = reinterpret_cast<const uint4*>(Kb + (int64_t)t * rs);
            const uint4* krb = reinterpret_cast<const uint4*>(Kb + (int64_t)(has1 ? t1 : t) * rs);
#pragma unroll
            for (int c8 = 0; c8 < 8; c8++) { ka[c8] = __ldg(kra + c8); kb[c8] = __ldg(krb + c8); }
        }
        float2 a0[kDtwpQB / 2], a1[kDtwpQB / 2];  // (query 2 qp, query 2 qp + 1) of row t / row t1
#pragma unroll
        for (int qp = 0; qp < kDtwpQB / 2; qp++) { a0[qp] = make_float2(0.0f, 0.0f); a1[qp] = make_float2(0.0f, 0.0f); }
#pragma unroll
        for (int c8 = 0; c8 < 8; c8++) {
            const __nv_bfloat162* ea = reinterpret_cast<const __nv_bfloat162*>(&ka[c8]);
            const __nv_bfloat162* eb = reinterpret_cast<const __nv_bfloat162*>(&kb[c8]);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float2 fa = __bfloat1622float2(ea[j]), fb = __bfloat1622float2(eb[j]);  // columns 2 c2, 2 c2 + 1 (c2 = 4 c8 + j)
#pragma unroll
                for (int qp = 0; qp < kDtwpQB / 2; qp++) {
                    const float4 f = *reinterpret_cast<const float4*>(qs + qp * 128 + (c8 * 4 + j) * 4);  // those columns of both queries
                    a0[qp] = __ffma2_rn(make_float2(f.x, f.y), make_float2(fa.x, fa.x), a0[qp]);
                    a0[qp] = __ffma2_rn(make_float2(f.z, f.w), make_float2(fa.y, fa.y), a0[qp]);
                    a1[qp] = __ffma2_rn(make_float2(f.x, f.y), make_float2(fb.x, fb.x), a1[qp]);
                    a1[qp] = __ffma2_rn(make_float2(f.z, f.w), make_float2(fb.y, fb.y), a1[qp]);
                }
            }
        }
#pragma unroll
        for (int qp = 0; qp < kDtwpQB / 2; qp++) {
            p[(2 * qp) * kDtwpPStride + t] = a0[qp].x;
            p[(2 * qp + 1) * kDtwpPStride + t] = a0[qp].y;
            if (has1) {
                p[(2 * qp) * kDtwpPStride + t1] = a1[qp].x;
                p[(2 * qp + 1) * kDtwpPStride + t1] = a1[qp].y;
            }
        }
    }
    __syncthreads();
    // ---- softmax (warp w: queries w, w + 8) and alignment-head capture ----
    const int ahead = ahead_map ? ahead_map[hh] : -1;
    for (int qi = warp; qi < nq; qi += 8) {
        float* pr = p + qi * kDtwpPStride;
        float mx = -INFINITY;
        for (int t = lane; t < kT; t += 32) mx = fmaxf(mx, pr[t]);
        mx = warp_max(mx);
        float sum = 0.0f;
        for (int t = lane; t < kT; t += 32) {
            const float e = expf(pr[t] - mx);
            pr[t] = e;
            sum += e;
        }
        sum = warp_sum(sum);
        const float inv = 1.0f / sum;
        for (int t = lane; t < kT; t += 32) pr[t] *= inv;
        if (ahead >= 0) {
            __syncwarp();
            const int A_b = aw_A[b];
            float* dst = aw + aw_off[b] + ((int64_t)ahead * T_b + (q0 + qi)) * A_b;
            for (int t = lane; t < A_b; t += 32) dst[t] = pr[t];
        }
    }
    __syncthreads();
    // ---- P V: warp = key slice (groups of 4 keys), lane = column pair ----
    float2 acc[kDtwpQB];  // packed FFMA2 with the probability broadcast: per column the same fmaf chain as a scalar loop
#pragma unroll
    for (int qi = 0; qi < kDtwpQB; qi++) acc[qi] = make_float2(0.0f, 0.0f);
    // two of the warp's key groups (t4 and t4 + 32) are loaded per trip and consumed in the same order as before: half as many
    // dependent load -> FFMA2 round trips, identical arithmetic
    for (int t4 = warp * 4; t4 < kT; t4 += 64) {
        const bool second = t4 + 32 < kT;  // warp-uniform
        float2 v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int t = t4 + (j < 4 ? j : 28 + j);
            v[j] = (j < 4 || second) ? __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(Vb + (int64_t)t * rs + 2 * lane)) : make_float2(0.0f, 0.0f);
        }
#pragma unroll
        for (int g2 = 0; g2 < 2; g2++) {
            if (g2 == 1 && !second) break;
#pragma unroll
            for (int qi = 0; qi < kDtwpQB; qi++) {
                const float4 pv = *reinterpret_cast<const float4*>(p + qi * kDtwpPStride + t4 + 32 * g2);
                acc[qi] = __ffma2_rn(make_float2(pv.x, pv.x), v[4 * g2], acc[qi]);
                acc[qi] = __ffma2_rn(make_float2(pv.y, pv.y), v[4 * g2 + 1], acc[qi]);
                acc[qi] = __ffma2_rn(make_float2(pv.z, pv.z), v[4 * g2 + 2], acc[qi]);
                acc[qi] = __ffma2_rn(make_float2(pv.w, pv.w), v[4 * g2 + 3], acc[qi]);
            }
        }
    }
    __syncthreads();  // all warps are done reading p
    float* red = p;   // [8][16][64]
#pragma unroll
    for (int qi = 0; qi < kDtwpQB; qi++) {
        red[(warp * kDtwpQB + qi) * 64 + 2 * lane] = acc[qi].x;
        red[(warp * kDtwpQB + qi) * 64 + 2 * lane + 1] = acc[qi].y;
    }
    __syncthreads();
    for (int e = tid; e < nq * 64; e += 256) {
        const int qi = e >> 6, c = e & 63;
        float a = 0.0f;
#pragma unroll
        for (int w = 0; w < 8; w++) a += red[(w * kDtwpQB + qi) * 64 + c];
        store_split(att, lo_off, (r0 + qi) * (int64_t)d + hh * 64 + c, a);
    }
}

// The same operator on the tensor cores (warp-level mma.sync m16n8k16, bf16 x bf16 -> fp32): 16 queries are exactly one M tile.
// K_c / V_c are bf16 already; the fp32 queries and probabilities enter as THREE bf16 terms each (x = hi + mid + lo, 24 mantissa
// bits, every product exact, fp32 accumulation), three MMAs per k-step on the same B fragment.  Fragments come straight from
// global memory with 16-byte loads — no shared-memory staging, no ldmatrix: the contraction index (scores: the 64 columns) and
// the output column index (P V) are PERMUTED so that the 2-element fragment entries a thread needs are adjacent in its 16 bytes:
//   scores, n-tile = 8 keys: thread (g = lane / 4, t = lane % 4) loads columns 8t .. 8t+7 and 32+8t .. 32+8t+7 of key g (each
//     warp instruction = eight 64-byte row halves); k-step s uses word pair (2s, 2s+1) of those 32 bytes, i.e. logical k = 2t+{0,1}
//     <-> column base(s) + 8t + {0,1}, logical k = 2t+8+{0,1} <-> column base(s) + 8t + 2 + {0,1}, base = {0, 4, 32, 36}; the
//     query fragments are built once per CTA with the same map;
//   P V, k-step = 16 keys: the thread loads columns 8g .. 8g+7 of keys 2t, 2t+1, 2t+8, 2t+9 (each warp instruction = four full
//     128-byte rows); word p of those loads holds physical columns 8g+2p, 8g+2p+1, which two byte permutes turn into the B fragments
//     of n-tiles 2p and 2p+1 — logical column n of n-tile 2p (2p+1) is physical column 8n + 2p (8n + 2p + 1).
// S stays fp32 in shared memory with the exact two-pass softmax of the kernel above (so the alignment heads' probabilities are
// produced by the same code); warp w takes n-tiles / k-steps w, w + 8, ..., partial O tiles are reduced through shared memory.
constexpr int kMqPitch = kT + 12;  // S row pitch: == 8 (mod 32), so the 64-bit fragment accesses of a warp spread over all banks
constexpr int kMqQPitch = 68;

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// (x0, x1) as three packed bf16 pairs: x = hi + mid + lo to 24 bits (x0 in the low half: the lower fragment index)
__device__ __forceinline__ void split3_bf16(float x0, float x1, uint32_t& hi, uint32_t& mid, uint32_t& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
    const float2 hf = __bfloat1622float2(h);
    const float r0 = x0 - hf.x, r1 = x1 - hf.y;
    const __nv_bfloat162 m = __floats2bfloat162_rn(r0, r1);
    const float2 mf = __bfloat1622float2(m);
    const __nv_bfloat162 l = __floats2bfloat162_rn(r0 - mf.x, r1 - mf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    mid = *reinterpret_cast<const uint32_t*>(&m);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

__global__ void __launch_bounds__(256, 2)
dtwp_cross_attn_mma_kernel(const float* __restrict__ qpart /* [M][d] */, const float* __restrict__ b_q, const __nv_bfloat16* __restrict__ ckv, int d,
                           const int32_t* __restrict__ row_off, const int32_t* __restrict__ T, __nv_bfloat16* __restrict__ att, int64_t lo_off,
                           const int32_t* __restrict__ ahead_map, float* __restrict__ aw, const int64_t* __restrict__ aw_off,
                           const int32_t* __restrict__ aw_A) {
    extern __shared__ float sm[];
    const int hh = blockIdx.y, b = blockIdx.z, q0 = blockIdx.x * kDtwpQB, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T_b = T[b];
    if (q0 >= T_b) return;
    const int nq = min(kDtwpQB, T_b - q0);
    const int64_t r0 = row_off[b] + q0;
    float* qs = sm;                          // [16][68], pre-scaled by 1/8; rows >= nq are zero
    float* p = qs + kDtwpQB * kMqQPitch;     // [16][kMqPitch]; reused as the P V reduction buffer [8][16][64]
    for (int e = tid; e < kDtwpQB * 64; e += 256) {
        const int qi = e >> 6, c = e & 63;
        qs[qi * kMqQPitch + c] = qi < nq ? (qpart[(r0 + qi) * (int64_t)d + hh * 64 + c] + b_q[hh * 64 + c]) * 0.125f : 0.0f;
    }
    for (int e = tid; e < kDtwpQB * (kMqPitch - kT); e += 256) p[(e / (kMqPitch - kT)) * kMqPitch + kT + e % (kMqPitch - kT)] = 0.0f;
    __syncthreads();
    const int g = lane >> 2, t = lane & 3;
    const __nv_bfloat16* Kb = ckv + ((int64_t)b * gridDim.y + hh) * 2 * kT * 64;
    const __nv_bfloat16* Vb = Kb + kT * 64;
    // ---- scores ----
    {
        uint32_t ah[4][4], am[4][4], al[4][4];  // query fragments of the four k-steps, three terms
#pragma unroll
        for (int s = 0; s < 4; s++) {
            const int c0 = (s & 1) * 4 + (s >> 1) * 32 + 8 * t;
            const float2 x0 = *reinterpret_cast<const float2*>(qs + g * kMqQPitch + c0);
            const float2 x1 = *reinterpret_cast<const float2*>(qs + (g + 8) * kMqQPitch + c0);
            const float2 x2 = *reinterpret_cast<const float2*>(qs + g * kMqQPitch + c0 + 2);
            const float2 x3 = *reinterpret_cast<const float2*>(qs + (g + 8) * kMqQPitch + c0 + 2);
            split3_bf16(x0.x, x0.y, ah[s][0], am[s][0], al[s][0]);
            split3_bf16(x1.x, x1.y, ah[s][1], am[s][1], al[s][1]);
            split3_bf16(x2.x, x2.y, ah[s][2], am[s][2], al[s][2]);
            split3_bf16(x3.x, x3.y, ah[s][3], am[s][3], al[s][3]);
        }
        constexpr int kNT = (kT + 7) / 8;  // 188 n-tiles; the last one holds 4 keys (the others are clamped re-reads, not stored)
        auto load_k = [&](int nt, uint4& u0, uint4& u1) {
            const uint4* src = reinterpret_cast<const uint4*>(Kb + (int64_t)min(nt * 8 + g, kT - 1) * 64 + 8 * t);
            u0 = __ldg(src);
            u1 = __ldg(src + 4);
        };
        uint4 u0, u1, v0, v1;
        load_k(warp, u0, u1);
        for (int nt = warp; nt < kNT; nt += 8) {
            const bool more = nt + 8 < kNT;  // warp-uniform
            if (more) load_k(nt + 8, v0, v1);
            float dd[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            mma_bf16_16816(dd, al[0], u0.x, u0.y); mma_bf16_16816(dd, am[0], u0.x, u0.y); mma_bf16_16816(dd, ah[0], u0.x, u0.y);
            mma_bf16_16816(dd, al[1], u0.z, u0.w); mma_bf16_16816(dd, am[1], u0.z, u0.w); mma_bf16_16816(dd, ah[1], u0.z, u0.w);
            mma_bf16_16816(dd, al[2], u1.x, u1.y); mma_bf16_16816(dd, am[2], u1.x, u1.y); mma_bf16_16816(dd, ah[2], u1.x, u1.y);
            mma_bf16_16816(dd, al[3], u1.z, u1.w); mma_bf16_16816(dd, am[3], u1.z, u1.w); mma_bf16_16816(dd, ah[3], u1.z, u1.w);
            const int k = nt * 8 + 2 * t;
            if (k < kT) {  // kT is even: k + 1 < kT as well
                *reinterpret_cast<float2*>(p + g * kMqPitch + k) = make_float2(dd[0], dd[1]);
                *reinterpret_cast<float2*>(p + (g + 8) * kMqPitch + k) = make_float2(dd[2], dd[3]);
            }
            if (more) { u0 = v0; u1 = v1; }
        }
    }
    __syncthreads();
    // ---- softmax (warp w: queries w, w + 8) and alignment-head capture ----
    const int ahead = ahead_map ? ahead_map[hh] : -1;
    for (int qi = warp; qi < nq; qi += 8) {
        float* pr = p + qi * kMqPitch;
        float mx = -INFINITY;
        for (int tt = lane; tt < kT; tt += 32) mx = fmaxf(mx, pr[tt]);
        mx = warp_max(mx);
        float sum = 0.0f;
        for (int tt = lane; tt < kT; tt += 32) {
            const float e = expf(pr[tt] - mx);
            pr[tt] = e;
            sum += e;
        }
        sum = warp_sum(sum);
        const float inv = 1.0f / sum;
        for (int tt = lane; tt < kT; tt += 32) pr[tt] *= inv;
        if (ahead >= 0) {
            __syncwarp();
            const int A_b = aw_A[b];
            float* dst = aw + aw_off[b] + ((int64_t)ahead * T_b + (q0 + qi)) * A_b;
            for (int tt = lane; tt < A_b; tt += 32) dst[tt] = pr[tt];
        }
    }
    __syncthreads();
    // ---- P V ----
    float o[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; nt++) { o[nt][0] = 0.0f; o[nt][1] = 0.0f; o[nt][2] = 0.0f; o[nt][3] = 0.0f; }
    {
        constexpr int kKS = (kT + 15) / 16;  // 94 k-steps; keys >= kT: clamped V rows against zero probabilities
        auto load_v = [&](int ks, uint4 (&w)[4]) {
            const int k0 = ks * 16 + 2 * t;
            w[0] = __ldg(reinterpret_cast<const uint4*>(Vb + (int64_t)min(k0, kT - 1) * 64 + 8 * g));
            w[1] = __ldg(reinterpret_cast<const uint4*>(Vb + (int64_t)min(k0 + 1, kT - 1) * 64 + 8 * g));
            w[2] = __ldg(reinterpret_cast<const uint4*>(Vb + (int64_t)min(k0 + 8, kT - 1) * 64 + 8 * g));
            w[3] = __ldg(reinterpret_cast<const uint4*>(Vb + (int64_t)min(k0 + 9, kT - 1) * 64 + 8 * g));
        };
        uint4 w[4], wn[4];
        load_v(warp, w);
        for (int ks = warp; ks < kKS; ks += 8) {
            const bool more = ks + 8 < kKS;  // warp-uniform
            if (more) load_v(ks + 8, wn);
            uint32_t ph[4], pm[4], pl[4];
            {
                const float* pp = p + ks * 16 + 2 * t;
                const float2 x0 = *reinterpret_cast<const float2*>(pp + g * kMqPitch);
                const float2 x1 = *reinterpret_cast<const float2*>(pp + (g + 8) * kMqPitch);
                const float2 x2 = *reinterpret_cast<const float2*>(pp + g * kMqPitch + 8);
                const float2 x3 = *reinterpret_cast<const float2*>(pp + (g + 8) * kMqPitch + 8);
                split3_bf16(x0.x, x0.y, ph[0], pm[0], pl[0]);
                split3_bf16(x1.x, x1.y, ph[1], pm[1], pl[1]);
                split3_bf16(x2.x, x2.y, ph[2], pm[2], pl[2]);
                split3_bf16(x3.x, x3.y, ph[3], pm[3], pl[3]);
            }
#pragma unroll
            for (int pp = 0; pp < 4; pp++) {
                const uint32_t w0 = reinterpret_cast<const uint32_t*>(&w[0])[pp], w1 = reinterpret_cast<const uint32_t*>(&w[1])[pp];
                const uint32_t w2 = reinterpret_cast<const uint32_t*>(&w[2])[pp], w3 = reinterpret_cast<const uint32_t*>(&w[3])[pp];
                const uint32_t e0 = __byte_perm(w0, w1, 0x5410), e1 = __byte_perm(w2, w3, 0x5410);  // physical column 8g + 2pp
                const uint32_t f0 = __byte_perm(w0, w1, 0x7632), f1 = __byte_perm(w2, w3, 0x7632);  // physical column 8g + 2pp + 1
                mma_bf16_16816(o[2 * pp], pl, e0, e1); mma_bf16_16816(o[2 * pp], pm, e0, e1); mma_bf16_16816(o[2 * pp], ph, e0, e1);
                mma_bf16_16816(o[2 * pp + 1], pl, f0, f1); mma_bf16_16816(o[2 * pp + 1], pm, f0, f1); mma_bf16_16816(o[2 * pp + 1], ph, f0, f1);
            }
            if (more) {
#pragma unroll
                for (int i = 0; i < 4; i++) w[i] = wn[i];
            }
        }
    }
    __syncthreads();  // all warps are done reading p
    float* red = p;   // [8][16][64]
#pragma unroll
    for (int pp = 0; pp < 4; pp++) {
        // n-tiles 2pp / 2pp+1, logical columns n = 2t, 2t+1 -> physical columns 8n + 2pp (+1): pairs of adjacent columns
        float* r = red + (warp * kDtwpQB) * 64 + 2 * pp;
        *reinterpret_cast<float2*>(r + g * 64 + 8 * (2 * t)) = make_float2(o[2 * pp][0], o[2 * pp + 1][0]);
        *reinterpret_cast<float2*>(r + g * 64 + 8 * (2 * t + 1)) = make_float2(o[2 * pp][1], o[2 * pp + 1][1]);
        *reinterpret_cast<float2*>(r + (g + 8) * 64 + 8 * (2 * t)) = make_float2(o[2 * pp][2], o[2 * pp + 1][2]);
        *reinterpret_cast<float2*>(r + (g + 8) * 64 + 8 * (2 * t + 1)) = make_float2(o[2 * pp][3], o[2 * pp + 1][3]);
    }
    __syncthreads();
    for (int e = tid; e < nq * 64; e += 256) {
        const int qi = e >> 6, c = e & 63;
        float a = 0.0f;
#pragma unroll
        for (int w = 0; w < 8; w++) a += red[(w * kDtwpQB + qi) * 64 + c];
        store_split(att, lo_off, (r0 + qi) * (int64_t)d + hh * 64 + c, a);
    }
}

// Beam search / best_of decoding on the same tensor-core scheme: the K <= 8 rows (decoders) of a window against its K_c / V_c, one CTA
// per (head, window), both 192 KB blocks read ONCE for all rows.  Eight queries fill half an M tile, so the three bf16 terms are
// laid out as A1 = [hi (rows 0-7); mid (rows 8-15)] and A2 = [lo; 0]: two MMAs per k-step into ONE accumulator, and the value of
// query g is c(row g) + c(row g + 8) — both live in the same thread.  Replaces dec_cross_attn_rows_kernel (fp32 FFMA2: 273 us per
// launch for 120 windows x 5 rows = 0.52 of the HBM roofline, the FMA pipe and its load -> FMA round trips in the way).
__global__ void __launch_bounds__(256, 2)
dec_cross_attn_rows_mma_kernel(const float* __restrict__ part, int n_splits, int64_t split_stride, const float* __restrict__ b_q,
                               const __nv_bfloat16* __restrict__ ckv, int d, __nv_bfloat16* __restrict__ att, int64_t lo_off, int pos,
                               const int32_t* __restrict__ t_limit, const int32_t* __restrict__ row_window, int K) {
    extern __shared__ float sm[];
    constexpr int NQ = 8;
    float* qs = sm;                     // [8][68], pre-scaled by 1/8; rows >= K are zero
    float* p = qs + NQ * kMqQPitch;     // [8][kMqPitch]; reused as the P V reduction buffer [8 warps][8][64]
    const int hh = blockIdx.x, w = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b0 = w * K;
    bool any = false;
    for (int i = 0; i < K; i++) any |= pos < t_limit[b0 + i];
    if (!any) return;
    for (int e = tid; e < NQ * 64; e += 256) {
        const int qi = e >> 6, c = e & 63;
        qs[qi * kMqQPitch + c] = qi < K ? (part_sum(part, n_splits, split_stride, (int64_t)(b0 + qi) * d + hh * 64 + c) + b_q[hh * 64 + c]) * 0.125f : 0.0f;
    }
    for (int e = tid; e < NQ * (kMqPitch - kT); e += 256) p[(e / (kMqPitch - kT)) * kMqPitch + kT + e % (kMqPitch - kT)] = 0.0f;
    __syncthreads();
    const int g = lane >> 2, t = lane & 3;
    const __nv_bfloat16* Kb = ckv + ((int64_t)row_window[b0] * gridDim.x + hh) * 2 * kT * 64;
    const __nv_bfloat16* Vb = Kb + kT * 64;
    // ---- scores ----
    {
        uint32_t a1[4][4], a2[4][4];
#pragma unroll
        for (int s = 0; s < 4; s++) {
            const int c0 = (s & 1) * 4 + (s >> 1) * 32 + 8 * t;
            const float2 x0 = *reinterpret_cast<const float2*>(qs + g * kMqQPitch + c0);
            const float2 x2 = *reinterpret_cast<const float2*>(qs + g * kMqQPitch + c0 + 2);
            split3_bf16(x0.x, x0.y, a1[s][0], a1[s][1], a2[s][0]);
            split3_bf16(x2.x, x2.y, a1[s][2], a1[s][3], a2[s][2]);
            a2[s][1] = 0u;
            a2[s][3] = 0u;
        }
        constexpr int kNT = (kT + 7) / 8;
        auto load_k = [&](int nt, uint4& u0, uint4& u1) {
            const uint4* src = reinterpret_cast<const uint4*>(Kb + (int64_t)min(nt * 8 + g, kT - 1) * 64 + 8 * t);
            u0 = __ldg(src);
            u1 = __ldg(src + 4);
        };
        // two n-tiles ahead: 3 x 1 KB per warp in flight (this kernel has to keep HBM busy, unlike the DTW-pass one)
        uint4 u0, u1, v0, v1, x0, x1;
        load_k(warp, u0, u1);
        load_k(min(warp + 8, kNT - 1), v0, v1);
        for (int nt = warp; nt < kNT; nt += 8) {
            load_k(min(nt + 16, kNT - 1), x0, x1);
            float dd[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            mma_bf16_16816(dd, a2[0], u0.x, u0.y); mma_bf16_16816(dd, a1[0], u0.x, u0.y);
            mma_bf16_16816(dd, a2[1], u0.z, u0.w); mma_bf16_16816(dd, a1[1], u0.z, u0.w);
            mma_bf16_16816(dd, a2[2], u1.x, u1.y); mma_bf16_16816(dd, a1[2], u1.x, u1.y);
            mma_bf16_16816(dd, a2[3], u1.z, u1.w); mma_bf16_16816(dd, a1[3], u1.z, u1.w);
            const int k = nt * 8 + 2 * t;
            if (k < kT) *reinterpret_cast<float2*>(p + g * kMqPitch + k) = make_float2(dd[0] + dd[2], dd[1] + dd[3]);
            u0 = v0; u1 = v1; v0 = x0; v1 = x1;
        }
    }
    __syncthreads();
    // ---- softmax: warp i = row i ----
    if (warp < K) {
        float* pr = p + warp * kMqPitch;
        float mx = -INFINITY;
        for (int tt = lane; tt < kT; tt += 32) mx = fmaxf(mx, pr[tt]);
        mx = warp_max(mx);
        float sum = 0.0f;
        for (int tt = lane; tt < kT; tt += 32) {
            const float e = expf(pr[tt] - mx);
            pr[tt] = e;
            sum += e;
        }
        sum = warp_sum(sum);
        const float inv = 1.0f / sum;
        for (int tt = lane; tt < kT; tt += 32) pr[tt] *= inv;
    }
    __syncthreads();
    // ---- P V ----
    float o[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; nt++) { o[nt][0] = 0.0f; o[nt][1] = 0.0f; o[nt][2] = 0.0f; o[nt][3] = 0.0f; }
    {
        constexpr int kKS = (kT + 15) / 16;
        auto load_v = [&](int ks, uint4 (&wv)[4]) {
            const int k0 = ks * 16 + 2 * t;
            wv[0] = __ldg(reinterpret_cast<const uint4*>(Vb + (int64_t)min(k0, kT - 1) * 64 + 8 * g));
            wv[1] = __ldg(reinterpret_cast<const uint4*>(Vb + (int64_t)min(k0 + 1, kT - 1) * 64 + 8 * g));
            wv[2] = __ldg(reinterpret_cast<const uint4*>(Vb + (int64_t)min(k0 + 8, kT - 1) * 64 + 8 * g));
            wv[3] = __ldg(reinterpret_cast<const uint4*>(Vb + (int64_t)min(k0 + 9, kT - 1) * 64 + 8 * g));
        };
        uint4 wc[4], wn[4];
        load_v(warp, wc);
        for (int ks = warp; ks < kKS; ks += 8) {
            load_v(min(ks + 8, kKS - 1), wn);
            uint32_t pa1[4], pa2[4];
            {
                const float* pp = p + g * kMqPitch + ks * 16 + 2 * t;
                const float2 x0 = *reinterpret_cast<const float2*>(pp);
                const float2 x2 = *reinterpret_cast<const float2*>(pp + 8);
                split3_bf16(x0.x, x0.y, pa1[0], pa1[1], pa2[0]);
                split3_bf16(x2.x, x2.y, pa1[2], pa1[3], pa2[2]);
                pa2[1] = 0u;
                pa2[3] = 0u;
            }
#pragma unroll
            for (int pp = 0; pp < 4; pp++) {
                const uint32_t w0 = reinterpret_cast<const uint32_t*>(&wc[0])[pp], w1 = reinterpret_cast<const uint32_t*>(&wc[1])[pp];
                const uint32_t w2 = reinterpret_cast<const uint32_t*>(&wc[2])[pp], w3 = reinterpret_cast<const uint32_t*>(&wc[3])[pp];
                const uint32_t e0 = __byte_perm(w0, w1, 0x5410), e1 = __byte_perm(w2, w3, 0x5410);
                const uint32_t f0 = __byte_perm(w0, w1, 0x7632), f1 = __byte_perm(w2, w3, 0x7632);
                mma_bf16_16816(o[2 * pp], pa2, e0, e1); mma_bf16_16816(o[2 * pp], pa1, e0, e1);
                mma_bf16_16816(o[2 * pp + 1], pa2, f0, f1); mma_bf16_16816(o[2 * pp + 1], pa1, f0, f1);
            }
#pragma unroll
            for (int i = 0; i < 4; i++) wc[i] = wn[i];
        }
    }
    __syncthreads();  // all warps are done reading p
    float* red = p;   // [8 warps][8][64]
#pragma unroll
    for (int pp = 0; pp < 4; pp++) {
        float* r = red + (warp * NQ + g) * 64 + 2 * pp;
        *reinterpret_cast<float2*>(r + 8 * (2 * t)) = make_float2(o[2 * pp][0] + o[2 * pp][2], o[2 * pp + 1][0] + o[2 * pp + 1][2]);
        *reinterpret_cast<float2*>(r + 8 * (2 * t + 1)) = make_float2(o[2 * pp][1] + o[2 * pp][3], o[2 * pp + 1][1] + o[2 * pp + 1][3]);
    }
    __syncthreads();
    for (int e = tid; e < K * 64; e += 256) {
        const int qi = e >> 6, c = e & 63;
        if (pos >= t_limit[b0 + qi]) continue;
        float a = 0.0f;
#pragma unroll
        for (int wv = 0; wv < 8; wv++) a += red[(wv * NQ + qi) * 64 + c];
        store_split(att, lo_off, (int64_t)(b0 + qi) * d + hh * 64 + c, a);
    }
}

static cudaError_t launch_cross_rows_mma(int H, int nW, int K, cudaStream_t st, const float* part, int n_splits, int64_t split_stride, const float* b_q,
                                         const __nv_bfloat16* ckv, int d, __nv_bfloat16* att, int64_t lo_off, int pos, const int32_t* t_limit,
                                         const int32_t* row_window) {
    const size_t smem = sizeof(float) * ((size_t)8 * kMqQPitch + (size_t)8 * kMqPitch);
    static DeviceOnce attr_once;
    {
        const cudaError_t e = per_device_once(attr_once, [&] { return cudaFuncSetAttribute(dec_cross_attn_rows_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); });
        if (e != cudaSuccess) return e;
    }
    dec_cross_attn_rows_mma_kernel<<<dim3(H, nW), 256, smem, st>>>(part, n_splits, split_stride, b_q, ckv, d, att, lo_off, pos, t_limit, row_window, K);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------
// whisper_process_logits + whisper_sample_token(best) + the decoder bookkeeping of whisper_full's inner loop
// ---------------------------------------------------------------------------------------------------
constexpr int kSampThreads = 1024;
constexpr int kSampPer = 51;  // 51 * 1024 = 52224 >= n_vocab

__global__ void __launch_bounds__(kSampThreads, 1)
dec_sample_kernel(const float* __restrict__ logits, int64_t ldv, DecWinState* __restrict__ win, wdr_token_data* __restrict__ tokens,
                  int32_t* __restrict__ seq, const int32_t* __restrict__ pos_ptr, int pos, const SampleParams sp, int32_t* __restrict__ done_count) {
    __shared__ float red[32];
    __shared__ unsigned long long red64[32];
    const int b = blockIdx.x, tid = threadIdx.x;
    pdl_launch_dependents();
    pdl_wait();
    pos = load_pos(pos_ptr, pos);
    DecWinState st = win[b];
    if (st.completed || st.failed) return;
    const int n = sp.n_vocab;
    const int n_cur = st.n_cur;
    const bool is_initial = n_cur == 0;
    const wdr_token_data* tk = tokens + (int64_t)b * kDecMaxTokens;
    const bool last_was_ts = n_cur > 0 && tk[n_cur - 1].id >= sp.beg;
    const bool penult_was_ts = n_cur < 2 || tk[n_cur - 2].id >= sp.beg;
    const float* lg = logits + (int64_t)b * ldv;
    float v[kSampPer];
#pragma unroll
    for (int k = 0; k < kSampPer; k++) {
        const int i = tid + k * kSampThreads;
        v[k] = (i < n) ? lg[i] : -INFINITY;
    }
    // no_speech_prob: softmax over the unfiltered logits, taken after the prompt (whisper_full)
    if (is_initial) {
        float m = -INFINITY;
#pragma unroll
        for (int k = 0; k < kSampPer; k++) m = fmaxf(m, v[k]);
        m = block_max(m, red);
        float s = 0.0f;
#pragma unroll
        for (int k = 0; k < kSampPer; k++) if (v[k] > -INFINITY) s += expf(v[k] - m);
        s = block_sum(s, red);
        if (tid == 0) st.no_speech_prob = expf(lg[sp.nosp] - (logf(s) + m));
    }
    // ---- logit rules ----
    const int ts_floor = st.has_ts ? sp.beg + st.seek_delta / 2 : sp.beg;
#pragma unroll
    for (int k = 0; k < kSampPer; k++) {
        const int i = tid + k * kSampThreads;
        bool mask = false;
        if (sp.suppress_blank && is_initial && (i == sp.eot || i == sp.space)) mask = true;
        if (i == sp.not_ || i == sp.sot || i == sp.nosp || i == sp.solm || i == sp.translate || i == sp.transcribe || i == sp.prev) mask = true;
        if (sp.no_timestamps && i >= sp.beg) mask = true;
        if (i >= sp.lang0 && i < sp.lang0 + sp.n_langs) mask = true;
        if (last_was_ts) {
            if (penult_was_ts) { if (i >= sp.beg) mask = true; }
            else { if (i < sp.eot) mask = true; }
        }
        if (is_initial && sp.initial_tid0 >= 0 && i >= sp.beg + sp.initial_tid0 + 1) mask = true;
        if (i >= sp.beg && i < ts_floor) mask = true;
        if (mask) v[k] = -INFINITY;
    }
    // ---- log-softmax ----
    float m = -INFINITY;
#pragma unroll
    for (int k = 0; k < kSampPer; k++) m = fmaxf(m, v[k]);
    m = block_max(m, red);
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < kSampPer; k++) if (v[k] > -INFINITY) s += expf(v[k] - m);
    s = block_sum(s, red);
    const float lse = logf(s) + m;
#pragma unroll
    for (int k = 0; k < kSampPer; k++) if (v[k] > -INFINITY) v[k] -= lse;  // v = logprobs
    // ---- if the probability mass on timestamps beats every text token, sample a timestamp ----
    float tmax = -INFINITY, xmax = -INFINITY;
#pragma unroll
    for (int k = 0; k < kSampPer; k++) {
        const int i = tid + k * kSampThreads;
        if (i >= sp.beg) tmax = fmaxf(tmax, v[k]); else xmax = fmaxf(xmax, v[k]);
    }
    tmax = block_max(tmax, red);
    xmax = block_max(xmax, red);
    float tsum = 0.0f;
#pragma unroll
    for (int k = 0; k < kSampPer; k++) {
        const int i = tid + k * kSampThreads;
        if (i >= sp.beg && v[k] > -INFINITY) tsum += expf(v[k] - tmax);
    }
    tsum = block_sum(tsum, red);
    const float ts_logprob = tsum > 0.0f ? logf(tsum) + tmax : -INFINITY;
    const bool mask_text = ts_logprob > xmax;
    // ---- probabilities, timestamp statistics, greedy argmax (first maximum wins) ----
    float sum_ts = 0.0f;
    unsigned long long best_ts = 0ull, best_all = 0ull;
#pragma unroll
    for (int k = 0; k < kSampPer; k++) {
        const int i = tid + k * kSampThreads;
        if (mask_text && i < sp.beg) v[k] = -INFINITY;
        const float pr = v[k] > -INFINITY ? expf(v[k]) : 0.0f;
        if (pr > 0.0f) {
            const unsigned long long key = ((unsigned long long)__float_as_uint(pr) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)i);
            if (i >= sp.beg) { sum_ts += pr; best_ts = key > best_ts ? key : best_ts; }
            best_all = key > best_all ? key : best_all;
        }
    }
    sum_ts = block_sum(sum_ts, red);
    best_ts = block_max_u64(best_ts, red64);
    best_all = block_max_u64(best_all, red64);
    if (tid != 0) return;
    wdr_token_data td;
    td.id = 0; td.tid = 0; td.p = 0.0f; td.plog = 0.0f; td.pt = 0.0f; td.ptsum = 0.0f; td.t0 = -1; td.t1 = -1; td.t_dtw = -1; td.vlen = 0.0f;
    {
        double max_ts = 0.0;
        if (best_ts) { td.tid = (int)(0xFFFFFFFFu - (unsigned)(best_ts & 0xFFFFFFFFull)); max_ts = (double)__uint_as_float((unsigned)(best_ts >> 32)); }
        td.pt = (float)(max_ts / ((double)sum_ts + 1e-10));
        td.ptsum = sum_ts;
    }
    if (best_all) {
        td.id = (int)(0xFFFFFFFFu - (unsigned)(best_all & 0xFFFFFFFFull));
        td.p = __uint_as_float((unsigned)(best_all >> 32));
        td.plog = logf(td.p);  // replaced below by the exact logprob
    }
    // plog is logprobs[id]; recompute it from the logit (thread 0 does not hold that element): logit - lse (masking never selects a masked id)
    td.plog = lg[td.id] - lse;
    if (td.id >= sp.beg) { td.tid = td.id; td.pt = td.p; }
    // ---- decoder bookkeeping (whisper_full inner loop, one decoder, T = 0) ----
    const int i = n_cur;
    tokens[(int64_t)b * kDecMaxTokens + n_cur] = td;
    st.n_cur = n_cur + 1;
    bool done = false;
    if (td.id > sp.beg) {
        const int sd_new = 2 * (td.id - sp.beg);
        if (st.has_ts && st.seek_delta > sd_new && st.result_len < i) { st.failed = 1; done = true; }
        else { st.seek_delta = sd_new; st.result_len = i + 1; st.has_ts = 1; }
    }
    if (!done && (td.id == sp.eot || (st.has_ts && st.seek + st.seek_delta + sp.delta_min >= st.seek_end))) {
        bool fail = false;
        if (st.result_len == 0 && !sp.no_timestamps) {
            if (st.seek + st.seek_delta + sp.delta_min >= st.seek_end) st.result_len = i + 1;
            else fail = true;
        }
        if (fail) st.failed = 1;
        else {
            if (sp.single_segment || sp.no_timestamps) { st.result_len = i + 1; st.seek_delta = 100 * 30; }
            st.completed = 1;
        }
        done = true;
    }
    if (!done && i == sp.n_max - 1 && (st.result_len == 0 || st.seek_delta < 100 * 30 / 2)) { st.failed = 1; done = true; }
    if (!done && pos + 1 < kDecSeqCap) seq[b * kDecSeqCap + pos + 1] = td.id;
    win[b] = st;
    if (done) atomicAdd(done_count, 1);
}

// Beam search: whisper_process_logits for one row from its own history summary (BeamRow), then the k best tokens by log-probability
// (ties: lower id) with the statistics whisper_sample_token_topk attaches to each (p, plog; tid, pt, ptsum of the distribution).
__global__ void __launch_bounds__(kSampThreads, 1)
dec_topk_kernel(const float* __restrict__ logits, int64_t ldv, const BeamRow* __restrict__ rows, const SampleParams sp, int k_top, float temperature,
                BeamCand* __restrict__ cands /* [rows][kBeamMax] */, float* __restrict__ no_speech /* [rows], written when n_cur == 0 */,
                float* __restrict__ probs_out /* optional [rows][ldv]: the processed LOG-probabilities (temperature sampling on the host); slot 1 of cands then carries the distribution's raw tid / pt / ptsum */) {
    __shared__ float red[32];
    __shared__ unsigned long long red64[32];
    const int b = blockIdx.x, tid = threadIdx.x;
    const BeamRow st = rows[b];
    if (!st.active) return;
    const int n = sp.n_vocab;
    const bool is_initial = st.n_cur == 0;
    const bool last_was_ts = st.n_cur > 0 && st.last_id >= sp.beg;
    const bool penult_was_ts = st.n_cur < 2 || st.penult_id >= sp.beg;
    const float* lg = logits + (int64_t)b * ldv;
    float v[kSampPer];
#pragma unroll
    for (int k = 0; k < kSampPer; k++) {
        const int i = tid + k * kSampThreads;
        v[k] = (i < n) ? lg[i] : -INFINITY;
    }
    // no_speech_prob: whisper_full takes it once, from the RAW logits after the prompt decode ("this has to be done before any logit
    // filtering") — before whisper_process_logits divides by the temperature
    if (is_initial) {
        float m = -INFINITY;
#pragma unroll
        for (int k = 0; k < kSampPer; k++) m = fmaxf(m, v[k]);
        m = block_max(m, red);
        float s = 0.0f;
#pragma unroll
        for (int k = 0; k < kSampPer; k++) if (v[k] > -INFINITY) s += expf(v[k] - m);
        s = block_sum(s, red);
        if (tid == 0) no_speech[b] = expf(lg[sp.nosp] - (logf(s) + m));
    }
    if (temperature > 0.0f) {  // whisper_process_logits: logits[i] /= temperature, first of all
#pragma unroll
        for (int k = 0; k < kSampPer; k++) v[k] = __fdiv_rn(v[k], temperature);
    }
    const int ts_floor = st.has_ts ? sp.beg + st.seek_delta / 2 : sp.beg;
#pragma unroll
    for (int k = 0; k < kSampPer; k++) {
        const int i = tid + k * kSampThreads;
        bool mask = false;
        if (sp.suppress_blank && is_initial && (i == sp.eot || i == sp.space)) mask = true;
        if (i == sp.not_ || i == sp.sot || i == sp.nosp || i == sp.solm || i == sp.translate || i == sp.transcribe || i == sp.prev) mask = true;
        if (sp.no_timestamps && i >= sp.beg) mask = true;
        if (i >= sp.lang0 && i < sp.lang0 + sp.n_langs) mask = true;
        if (last_was_ts) {
            if (penult_was_ts) { if (i >= sp.beg) mask = true; }
            else { if (i < sp.eot) mask = true; }
        }
        if (is_initial && sp.initial_tid0 >= 0 && i >= sp.beg + sp.initial_tid0 + 1) mask = true;
        if (i >= sp.beg && i < ts_floor) mask = true;
        if (mask) v[k] = -INFINITY;
    }
    float m = -INFINITY;
#pragma unroll
    for (int k = 0; k < kSampPer; k++) m = fmaxf(m, v[k]);
    m = block_max(m, red);
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < kSampPer; k++) if (v[k] > -INFINITY) s += expf(v[k] - m);
    s = block_sum(s, red);
    const float lse = logf(s) + m;
#pragma unroll
    for (int k = 0; k < kSampPer; k++) if (v[k] > -INFINITY) v[k] -= lse;  // v = logprobs
    float tmax = -INFINITY, xmax = -INFINITY;
#pragma unroll
    for (int k = 0; k < kSampPer; k++) {
        const int i = tid + k * kSampThreads;
        if (i >= sp.beg) tmax = fmaxf(tmax, v[k]); else xmax = fmaxf(xmax, v[k]);
    }
    tmax = block_max(tmax, red);
    xmax = block_max(xmax, red);
    float tsum = 0.0f;
#pragma unroll
    for (int k = 0; k < kSampPer; k++) {
        const int i = tid + k * kSampThreads;
        if (i >= sp.beg && v[k] > -INFINITY) tsum += expf(v[k] - tmax);
    }
    tsum = block_sum(tsum, red);
    const float ts_logprob = tsum > 0.0f ? logf(tsum) + tmax : -INFINITY;
    const bool mask_text = ts_logprob > xmax;
    float sum_ts = 0.0f;
    unsigned long long best_ts = 0ull;
#pragma unroll
    for (int k = 0; k < kSampPer; k++) {
        const int i = tid + k * kSampThreads;
        if (mask_text && i < sp.beg) v[k] = -INFINITY;
        const float pr = v[k] > -INFINITY ? expf(v[k]) : 0.0f;
        if (probs_out && i < n) probs_out[(int64_t)b * ldv + i] = v[k];  // processed log-probability (-inf = masked)
        if (pr > 0.0f && i >= sp.beg) {
            const unsigned long long key = ((unsigned long long)__float_as_uint(pr) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)i);
            sum_ts += pr;
            best_ts = key > best_ts ? key : best_ts;
        }
    }
    sum_ts = block_sum(sum_ts, red);
    best_ts = block_max_u64(best_ts, red64);
    int tid_tok = 0;
    float pt = 0.0f;
    {
        double max_ts = 0.0;
        if (best_ts) { tid_tok = (int)(0xFFFFFFFFu - (unsigned)(best_ts & 0xFFFFFFFFull)); max_ts = (double)__uint_as_float((unsigned)(best_ts >> 32)); }
        pt = (float)(max_ts / ((double)sum_ts + 1e-10));
    }
    // ---- k rounds of block arg-max over the log-probabilities (key = order-preserving float bits | inverted id) ----
    for (int r = 0; r < k_top; r++) {
        unsigned long long best = 0ull;
#pragma unroll
        for (int k = 0; k < kSampPer; k++) {
            const int i = tid + k * kSampThreads;
            if (v[k] > -INFINITY) {
                const unsigned long long key = ((unsigned long long)float_to_key(v[k]) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)i);
                best = key > best ? key : best;
            }
        }
        best = block_max_u64(best, red64);
        BeamCand c;
        c.id = -1; c.tid = tid_tok; c.p = 0.0f; c.plog = -INFINITY; c.pt = pt; c.ptsum = sum_ts;
        if (best) {
            const int id = (int)(0xFFFFFFFFu - (unsigned)(best & 0xFFFFFFFFull));
            const float lp = key_to_float((unsigned)(best >> 32));
            c.id = id; c.plog = lp; c.p = expf(lp);
            if (id >= sp.beg) { c.tid = id; c.pt = c.p; }
#pragma unroll
            for (int k = 0; k < kSampPer; k++)
                if (tid + k * kSampThreads == id) v[k] = -INFINITY;  // taken
        }
        if (tid == 0) cands[(int64_t)b * kBeamMax + r] = c;
        __syncthreads();
    }
    if (probs_out && tid == 0 && k_top < kBeamMax) {  // what whisper_sample_token attaches to a drawn token before its own override
        BeamCand c;
        c.id = -1; c.tid = tid_tok; c.p = 0.0f; c.plog = -INFINITY; c.pt = pt; c.ptsum = sum_ts;
        cands[(int64_t)b * kBeamMax + k_top] = c;
    }
}

// ancestry update after a beam reassignment at sampling position i (0-based): new row b continues old row parent[b];
// the history position that old row just wrote (pos_last) now belongs to parent[b], earlier ones follow the parent's ancestry.
// Tables are row-major [row][ld = 448]: one CTA per new row copies its parent's history (coalesced) and appends the parent itself.
__global__ void beam_anc_kernel(const int32_t* __restrict__ anc_old, int32_t* __restrict__ anc_new, const int32_t* __restrict__ parent, int n_rows, int ld,
                                int pos_last) {
    const int b = blockIdx.x;
    if (b >= n_rows) return;
    const int pb = parent[b];
    for (int t = threadIdx.x; t < pos_last; t += blockDim.x) anc_new[(int64_t)b * ld + t] = anc_old[(int64_t)pb * ld + t];
    if (threadIdx.x == 0) anc_new[(int64_t)b * ld + pos_last] = pb;
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
void DecoderWorkspace::release() {
    for (void* p : {(void*)enc_bf16, (void*)sk, (void*)sv, (void*)x, (void*)h, (void*)att, (void*)ff, (void*)part, (void*)logits, (void*)seq,
                    (void*)tokens, (void*)win, (void*)done_count, (void*)pos_dev, (void*)beam_anc[0], (void*)beam_anc[1], (void*)beam_limit, (void*)beam_rows,
                    (void*)beam_cands, (void*)beam_parent, (void*)beam_nosp, (void*)beam_rowwin, (void*)ahead_map, (void*)aw, (void*)aw_off, (void*)aw_T, (void*)aw_A, (void*)cross_stats})
        if (p) cudaFree(p);
    for (auto p : ckv) if (p) cudaFree(p);
    if (step_graph) cudaGraphExecDestroy(step_graph);
    *this = DecoderWorkspace();
}

int DecoderWorkspace::reserve(const wdr_context* ctx, int windows, int rows) {
    if (rows < windows) rows = windows;
    if (windows <= cap_W && rows <= cap_B) return WDR_OK;
    WDR_REQUIRE(windows <= kDecMaxWindows && rows <= kDecMaxRows, "decode batch exceeds 128 windows / 640 rows");
    windows = std::max(windows, cap_W);
    rows = std::max(rows, cap_B);
    release();
    const int W = windows, B = rows;
    const WhisperArch& a = ctx->arch;
    d = a.d; n_layer = a.n_dec_layer; n_head = a.n_head;
    ldv = (a.n_vocab + 7) / 8 * 8;
    WDR_CUDA_TRY(cudaMalloc(&enc_bf16, sizeof(__nv_bfloat16) * (size_t)W * kT * d));
    ckv.assign(n_layer, nullptr);
    for (int l = 0; l < n_layer; l++) WDR_CUDA_TRY(cudaMalloc(&ckv[l], sizeof(__nv_bfloat16) * (size_t)W * kT * 2 * d));
    const size_t skv = (size_t)n_layer * B * kDecSeqCap * d;
    WDR_CUDA_TRY(cudaMalloc(&sk, sizeof(__half) * skv));
    WDR_CUDA_TRY(cudaMalloc(&sv, sizeof(__half) * skv));
    WDR_CUDA_TRY(cudaMalloc(&x, sizeof(float) * (size_t)B * d));
    WDR_CUDA_TRY(cudaMalloc(&h, sizeof(__nv_bfloat16) * (size_t)2 * B * d));    // (hi, lo) planes
    WDR_CUDA_TRY(cudaMalloc(&att, sizeof(__nv_bfloat16) * (size_t)2 * B * d));
    WDR_CUDA_TRY(cudaMalloc(&ff, sizeof(__nv_bfloat16) * (size_t)2 * B * 4 * d));
    part_elems = (size_t)32 * B * 4 * d;
    WDR_CUDA_TRY(cudaMalloc(&part, sizeof(float) * part_elems));
    WDR_CUDA_TRY(cudaMalloc(&logits, sizeof(float) * (size_t)B * ldv));
    WDR_CUDA_TRY(cudaMalloc(&seq, sizeof(int32_t) * (size_t)B * kDecSeqCap));
    WDR_CUDA_TRY(cudaMalloc(&tokens, sizeof(wdr_token_data) * (size_t)B * kDecMaxTokens));
    WDR_CUDA_TRY(cudaMalloc(&win, sizeof(DecWinState) * B));
    WDR_CUDA_TRY(cudaMalloc(&done_count, sizeof(int32_t)));
    WDR_CUDA_TRY(cudaMalloc(&pos_dev, sizeof(int32_t)));
    for (int i = 0; i < 2; i++) WDR_CUDA_TRY(cudaMalloc(&beam_anc[i], sizeof(int32_t) * kDecSeqCap * kDecMaxBatch));
    WDR_CUDA_TRY(cudaMalloc(&beam_limit, sizeof(int32_t) * kDecMaxBatch));
    WDR_CUDA_TRY(cudaMalloc(&beam_rows, sizeof(BeamRow) * kDecMaxBatch));
    WDR_CUDA_TRY(cudaMalloc(&beam_cands, sizeof(BeamCand) * kDecMaxBatch * kBeamMax));
    WDR_CUDA_TRY(cudaMalloc(&beam_parent, sizeof(int32_t) * kDecMaxBatch));
    WDR_CUDA_TRY(cudaMalloc(&beam_nosp, sizeof(float) * kDecMaxBatch));
    WDR_CUDA_TRY(cudaMalloc(&beam_rowwin, sizeof(int32_t) * kDecMaxBatch));
    WDR_CUDA_TRY(cudaMalloc(&ahead_map, sizeof(int32_t) * n_layer * n_head));
    WDR_CUDA_TRY(cudaMalloc(&aw_off, sizeof(int64_t) * B));
    WDR_CUDA_TRY(cudaMalloc(&aw_T, sizeof(int32_t) * B));
    WDR_CUDA_TRY(cudaMalloc(&aw_A, sizeof(int32_t) * B));
    WDR_CUDA_TRY(cudaMalloc(&cross_stats, sizeof(unsigned long long) * 2));
    WDR_CUDA_TRY(cudaMemset(cross_stats, 0, sizeof(unsigned long long) * 2));
    {
        std::vector<int32_t> map((size_t)n_layer * n_head, -1);
        n_aheads = (int)ctx->aheads.size();
        for (int i = 0; i < n_aheads; i++) {
            const int l = ctx->aheads[i].first, hd = ctx->aheads[i].second;
            if (l < n_layer && hd < n_head) map[(size_t)l * n_head + hd] = i;
        }
        WDR_CUDA_TRY(cudaMemcpy(ahead_map, map.data(), sizeof(int32_t) * map.size(), cudaMemcpyHostToDevice));
    }
    cap_B = B;
    cap_W = W;
    return WDR_OK;
}

int decoder_cross_kv(const wdr_context* ctx, DecoderWorkspace& ws, int B, cudaStream_t st, Profiler* prof) {
    const int d = ctx->arch.d;
    for (int l = 0; l < ctx->arch.n_dec_layer; l++) {
        const DecLayerW& e = ctx->w.dec[l];
        GemmDesc g;
        g.A = ws.enc_bf16; g.a_row_stride = d; g.rows_per_batch = B * kT; g.n_batch = 1;
        g.W = e.w_ckv; g.ldw = d; g.N = 2 * d; g.K = d;
        g.epilogue = EPI_HEADS_BF16; g.out = ws.ckv[l]; g.ldc = 2 * d; g.bias = e.b_ckv; g.group_rows = kT;
        ProfScope ps(prof, KC_GEMM, st);
        int rc = gemm_bf16(g, st);
        if (rc != WDR_OK) return rc;
    }
    return WDR_OK;
}

// Tile width and split-K factor of a weight-streaming GEMM (M = one 128-row tile).  Every work item loads its K-slice of the
// (hi, lo) activations (512 B per k), its weight tile (2*bn B per k) and writes a 128 x bn fp32 partial; measured on B200
// (ncu, in-graph): item time ~ fixed + bytes / ~100 GB/s per SM, and a second wave costs a whole item time again.  So: one wave
// (items <= SMs), the smallest per-item byte count, a small penalty per split for the consumer's partial-sum reads.
static void pick_tile(int K, int N, int rows, int* bn_out, int* splits_out) {
    const int num_kb = (K + 63) / 64, sms = 148, m_tiles = (rows + 127) / 128;  // beam batches: several 128-row M tiles
    static const double c_fixed = getenv("WDR_PT_FIXED") ? atof(getenv("WDR_PT_FIXED")) : 300.0;   // tuning knobs of the cost model below
    static const double c_split = getenv("WDR_PT_SPLIT") ? atof(getenv("WDR_PT_SPLIT")) : 12.0;
    static const double c_bytes = getenv("WDR_PT_BYTES") ? atof(getenv("WDR_PT_BYTES")) : 1.0;
    double best = 1e30;
    int best_bn = 64, best_s = 1;
    for (int bn : {64, 128}) {
        const int tiles = (N + bn - 1) / bn;
        for (int s = 1; s <= 16 && s <= num_kb; s++) {
            const int per = (num_kb + s - 1) / s;
            if ((num_kb + per - 1) / per != s) continue;  // every split must own >= 1 k-block
            const int items = tiles * s * m_tiles, waves = (items + sms - 1) / sms;
            const double ks = per * 64.0;
            const double kb = (512.0 * ks + 2.0 * bn * ks + 512.0 * bn) / 1024.0;
            const double cost = waves * (c_fixed + c_bytes * kb) + c_split * s;
            if (cost < best) { best = cost; best_bn = bn; best_s = s; }
        }
    }
    *bn_out = best_bn;
    *splits_out = best_s;
}

struct SkinnyGemm {
    int splits;
    int64_t split_stride;
};

// part[s][B][N] = A[B][K] * W[N][K]^T over K-slice s
static int skinny_gemm(const __nv_bfloat16* A, int B, const __nv_bfloat16* W, int N, int K, DecoderWorkspace& ws, SkinnyGemm* out,
                       cudaStream_t st, Profiler* prof, bool pdl) {
    GemmDesc g;
    g.A = A; g.a_row_stride = K; g.rows_per_batch = B; g.n_batch = 1;
    g.W = W; g.ldw = K; g.N = N; g.K = K;
    g.epilogue = EPI_F32; g.out = ws.part; g.ldc = N;
    g.dual_a = true; g.a_dual_stride = (int64_t)ws.cap_B * K;
    pick_tile(K, N, B, &g.bn, &g.split_k);
    g.pdl = pdl;
    g.w_kb_major = true;
    g.split_stride = (int64_t)B * N;
    if ((size_t)g.split_k * B * N > ws.part_elems) { set_error("decoder: split-K workspace too small"); return WDR_ERR_INVALID; }
    out->splits = g.split_k;
    out->split_stride = g.split_stride;
    ProfScope ps(prof, KC_DEC_GEMM, st);
    return gemm_bf16(g, st);
}

int decoder_step(const wdr_context* ctx, DecoderWorkspace& ws, int B, int pos, bool want_logits, int mode, cudaStream_t st, Profiler* prof,
                 bool pos_on_device, bool pdl) {
    const int32_t* pos_ptr = pos_on_device ? ws.pos_dev : nullptr;
    const WhisperArch& a = ctx->arch;
    const WhisperWeights& w = ctx->w;
    const int d = a.d, H = a.n_head, L = a.n_dec_layer;
    WDR_REQUIRE(pos >= 0 && pos < kDecSeqCap && B > 0 && B <= ws.cap_B, "decoder_step: bad position or batch");
    int rc;
    SkinnyGemm sg{1, 0};
    const float* pend_bias = nullptr;  // bias of the GEMM whose partials the next dec_ln folds into x
    bool pending = false;
    {
        ProfScope ps(prof, KC_DECODER, st);
        WDR_CUDA_TRY(launch_kernel(dec_embed_kernel, dim3(B), dim3(128), 0, st, pdl, ws.seq, pos_ptr, pos, w.tok_emb, w.dec_pos, d, a.n_vocab, ws.x));
        WDR_LAUNCH_CHECK();
    }
    // measurement aid (tools/full_phases.py): leave kernels of the chain out to see what each costs INSIDE the replayed graph — results
    // are garbage.  bit 0 self-attention, 1 LayerNorms, 2 cross-attention, 3 QKV, 4 out-projection, 5 cross-query, 6 cross-out, 7 fc1, 8 fc2
    static const int dbg_skip = getenv("WDR_DEBUG_SKIP") ? atoi(getenv("WDR_DEBUG_SKIP")) : 0;
    auto ln = [&](const float* g, const float* b) -> int {
        if (dbg_skip & 2) { pending = false; return WDR_OK; }
        ProfScope ps(prof, KC_DECODER, st);
        WDR_CUDA_TRY(launch_kernel(dec_ln_kernel, dim3(B), dim3(d / 4), 0, st, pdl, ws.x, pending ? ws.part : nullptr, sg.splits, sg.split_stride, d, pend_bias, g, b,
                                   ws.h, (int64_t)ws.cap_B * d, d));
        WDR_LAUNCH_CHECK();
        pending = false;
        return WDR_OK;
    };
    static const int dbg_layers = getenv("WDR_DEBUG_DEC_LAYERS") ? atoi(getenv("WDR_DEBUG_DEC_LAYERS")) : 1 << 30;  // bring-up aid: truncate the stack
    const bool capture = mode == DEC_MODE_DTW;
    const bool beam = mode == DEC_MODE_BEAM;  // rows = windows x beams: shared cross cache per window, ancestry-indexed self cache
    const DecWinState* win = mode == DEC_MODE_DECODE ? ws.win : nullptr;
    const int32_t* t_limit = mode == DEC_MODE_DTW ? ws.aw_T : beam ? ws.beam_limit : nullptr;
    int L_run = L;
    if (capture && !want_logits) {  // the DTW pass only needs the layers up to the last alignment head
        L_run = 0;
        for (auto& lh : ctx->aheads) L_run = std::max(L_run, lh.first + 1);
    }
    for (int l = 0; l < L_run && l < dbg_layers; l++) {
        const DecLayerW& e = w.dec[l];
        if ((rc = ln(e.ln1_g, e.ln1_b)) != WDR_OK) return rc;
        if (!(dbg_skip & 8) && (rc = skinny_gemm(ws.h, B, e.w_qkv, 3 * d, d, ws, &sg, st, prof, pdl)) != WDR_OK) return rc;
        if (!(dbg_skip & 1)) {
            ProfScope ps(prof, KC_DECODER, st);
            int hpc = kSelfMaxHeadsPerCta;
            while (H % hpc) hpc--;
            // (measured and dropped: one CTA per (window, head) over its K beam rows, so that siblings' shared history hits L1 — 748 vs 766
            // audio-s/s at beam 5, and the wider launch bounds / shared arrays cost the greedy instantiation 125 ms per step)
            WDR_CUDA_TRY(launch_kernel(beam ? dec_self_attn_kernel<true> : dec_self_attn_kernel<false>, dim3(H / hpc, B), dim3(hpc * 32), 0, st, pdl, ws.part,
                                       sg.splits, sg.split_stride, e.b_qkv, ws.sk + (size_t)l * ws.cap_B * kDecSeqCap * d,
                                       ws.sv + (size_t)l * ws.cap_B * kDecSeqCap * d, pos_ptr, pos, d, ws.att, (int64_t)ws.cap_B * d, win, t_limit,
                                       beam ? (const int32_t*)ws.beam_anc_cur : (const int32_t*)nullptr, kDecSeqCap));
            WDR_LAUNCH_CHECK();
        }
        if (!(dbg_skip & 16) && (rc = skinny_gemm(ws.att, B, e.w_o, d, d, ws, &sg, st, prof, pdl)) != WDR_OK) return rc;
        pending = true; pend_bias = e.b_o;
        if ((rc = ln(e.ln2_g, e.ln2_b)) != WDR_OK) return rc;
        if (!(dbg_skip & 32) && (rc = skinny_gemm(ws.h, B, e.w_cq, d, d, ws, &sg, st, prof, pdl)) != WDR_OK) return rc;
        if (!(dbg_skip & 4)) {
            ProfScope ps(prof, KC_DEC_CROSS, st);
            // register budget capped for 6 resident CTAs per SM (40 registers): measured 1900 ms of decode vs 1930 ms at 5 (48
            // registers); 8 (32 registers) spills and is much slower.  A single-pass online-softmax variant that streams K_c and
            // V_c rows together was tried and is slower (2070 ms): its per-iteration max -> exp -> FMA chain keeps fewer loads in
            // flight than the two independent passes do.
            static const bool no_rows = getenv("WDR_NO_ROWS_KERNEL") != nullptr;  // bring-up aid: one CTA per row even when rows share a window
            if (beam && ws.beam_K > 1 && B % ws.beam_K == 0 && !pos_on_device && !no_rows) {
                const int nW = B / ws.beam_K, nqp = (ws.beam_K + 1) / 2;
                cudaError_t ce;
#define WDR_ROWS(N) launch_cross_rows<N>(H, nW, ws.beam_K, st, ws.part, sg.splits, sg.split_stride, e.b_cq, ws.ckv[l], d, ws.att, (int64_t)ws.cap_B * d, pos, t_limit, ws.beam_rowwin)
                static const bool rows_ffma = getenv("WDR_ROWS_FFMA") != nullptr;  // A/B knob: the fp32-FMA rows kernel
                if (!rows_ffma && ws.beam_K <= 8)
                    ce = launch_cross_rows_mma(H, nW, ws.beam_K, st, ws.part, sg.splits, sg.split_stride, e.b_cq, ws.ckv[l], d, ws.att, (int64_t)ws.cap_B * d, pos, t_limit, ws.beam_rowwin);
                else
                ce = nqp == 1 ? WDR_ROWS(1) : nqp == 2 ? WDR_ROWS(2) : nqp == 3 ? WDR_ROWS(3) : WDR_ROWS(4);
#undef WDR_ROWS
                WDR_CUDA_TRY(ce);
            } else
            WDR_CUDA_TRY(launch_kernel(beam ? dec_cross_attn_kernel<6, true> : dec_cross_attn_kernel<6, false>, dim3(H, B), dim3(256), 0, st, pdl, ws.part, sg.splits, sg.split_stride, e.b_cq, ws.ckv[l], d, ws.att,
                                       (int64_t)ws.cap_B * d, capture ? ws.ahead_map + (size_t)l * H : nullptr, ws.aw, ws.aw_off, ws.aw_T, ws.aw_A, pos_ptr, pos,
                                       win, t_limit, beam ? (const int32_t*)ws.beam_rowwin : (const int32_t*)nullptr, ws.cross_stats));
            WDR_LAUNCH_CHECK();
        }
        if (!(dbg_skip & 64) && (rc = skinny_gemm(ws.att, B, e.w_co, d, d, ws, &sg, st, prof, pdl)) != WDR_OK) return rc;
        pending = true; pend_bias = e.b_co;
        if ((rc = ln(e.ln3_g, e.ln3_b)) != WDR_OK) return rc;
        if (!(dbg_skip & 128)) {   // fc1 with the bias + GELU + (hi, lo) split fused into the epilogue (N = 4d gives >= 24 x 64-wide tiles; no split-K)
            GemmDesc g;
            g.A = ws.h; g.a_row_stride = d; g.rows_per_batch = B; g.n_batch = 1;
            g.W = e.w_fc1; g.ldw = d; g.N = 4 * d; g.K = d;
            g.epilogue = EPI_BIAS_GELU_SPLIT; g.out = ws.ff; g.ldc = 4 * d; g.bias = e.b_fc1; g.bn = 64;
            g.dual_a = true; g.a_dual_stride = (int64_t)ws.cap_B * d;
            g.split_stride = (int64_t)ws.cap_B * 4 * d;
            g.pdl = pdl;
            g.w_kb_major = true;
            ProfScope ps(prof, KC_DEC_GEMM, st);
            if ((rc = gemm_bf16(g, st)) != WDR_OK) return rc;
        }
        if (!(dbg_skip & 256) && (rc = skinny_gemm(ws.ff, B, e.w_fc2, d, 4 * d, ws, &sg, st, prof, pdl)) != WDR_OK) return rc;
        pending = true; pend_bias = e.b_fc2;
    }
    if (want_logits) {
        if ((rc = ln(w.dec_ln_g, w.dec_ln_b)) != WDR_OK) return rc;
        GemmDesc g;
        g.A = ws.h; g.a_row_stride = d; g.rows_per_batch = B; g.n_batch = 1;
        g.W = w.tok_emb; g.ldw = d; g.N = (int)ws.ldv; g.K = d;
        // Tile width of the logits GEMM (51 866 x d weights = 133 MB, every item re-reads its (hi, lo) activation slice from L2): measured per
        // launch at 120 windows (ncu, in-graph) 56.1 us with 64-wide tiles (811 items), 41.8 us with 128 (406 items: half the activation traffic),
        // 40.7 us with 256 (203 items = 1.4 waves; instantiation not kept)
        static const int logits_bn = getenv("WDR_LOGITS_BN") ? atoi(getenv("WDR_LOGITS_BN")) : 128;
        g.epilogue = EPI_F32; g.out = ws.logits; g.ldc = ws.ldv; g.bn = logits_bn == 64 ? 64 : 128;
        g.dual_a = true; g.a_dual_stride = (int64_t)ws.cap_B * d;
        g.pdl = pdl;
        ProfScope ps(prof, KC_DEC_GEMM, st);
        if ((rc = gemm_bf16(g, st)) != WDR_OK) return rc;
    }
    return WDR_OK;
}

void DtwPassWorkspace::release() {
    for (void* p : {(void*)x, (void*)h, (void*)att, (void*)ff, (void*)part, (void*)row_b, (void*)row_pos, (void*)row_off})
        if (p) cudaFree(p);
    *this = DtwPassWorkspace();
}

int DtwPassWorkspace::reserve(int64_t rows, int d_model) {
    if (rows <= cap_rows && d_model == d) return WDR_OK;
    release();
    d = d_model;
    const size_t M = (size_t)((rows + 127) / 128 * 128);
    WDR_CUDA_TRY(cudaMalloc(&x, sizeof(float) * M * d));
    WDR_CUDA_TRY(cudaMalloc(&h, sizeof(__nv_bfloat16) * 2 * M * d));
    WDR_CUDA_TRY(cudaMalloc(&att, sizeof(__nv_bfloat16) * 2 * M * d));
    WDR_CUDA_TRY(cudaMalloc(&ff, sizeof(__nv_bfloat16) * 2 * M * 4 * d));
    WDR_CUDA_TRY(cudaMalloc(&part, sizeof(float) * M * 4 * d));
    WDR_CUDA_TRY(cudaMalloc(&row_b, sizeof(int32_t) * M));
    WDR_CUDA_TRY(cudaMalloc(&row_pos, sizeof(int32_t) * M));
    WDR_CUDA_TRY(cudaMalloc(&row_off, sizeof(int32_t) * kDecMaxWindows));
    cap_rows = (int64_t)M;
    return WDR_OK;
}

// out[M][N] (fp32) = (hi, lo)[M][K] * W[N][K]^T, full-size tcgen05 GEMM over the packed rows
static int packed_gemm(const __nv_bfloat16* A, int64_t M, int64_t M_cap, const __nv_bfloat16* W, int N, int K, float* out, cudaStream_t st, Profiler* prof) {
    GemmDesc g;
    g.A = A; g.a_row_stride = K; g.rows_per_batch = (int)M; g.n_batch = 1;
    g.W = W; g.ldw = K; g.N = N; g.K = K;
    g.epilogue = EPI_F32; g.out = out; g.ldc = N; g.bn = 128;
    g.dual_a = true; g.a_dual_stride = M_cap * K;
    g.w_kb_major = true;
    ProfScope ps(prof, KC_DEC_GEMM, st);
    return gemm_bf16(g, st);
}

int decoder_dtw_pass(const wdr_context* ctx, DecoderWorkspace& ws, DtwPassWorkspace& pw, int B, const int32_t* T_host, cudaStream_t st, Profiler* prof) {
    const WhisperArch& a = ctx->arch;
    const WhisperWeights& w = ctx->w;
    const int d = a.d, H = a.n_head;
    WDR_REQUIRE(B > 0 && B <= ws.cap_W && B <= kDecMaxWindows, "decoder_dtw_pass: bad batch");
    std::vector<int32_t> row_b, row_pos, row_off(kDecMaxWindows, 0);
    int max_T = 0;
    for (int b = 0; b < B; b++) {
        WDR_REQUIRE(T_host[b] >= 0 && T_host[b] <= kDtwpMaxT && T_host[b] <= kDecSeqCap, "decoder_dtw_pass: sequence too long");
        row_off[b] = (int32_t)row_b.size();
        for (int i = 0; i < T_host[b]; i++) { row_b.push_back(b); row_pos.push_back(i); }
        max_T = std::max(max_T, T_host[b]);
    }
    const int64_t M = (int64_t)row_b.size();
    if (M == 0) return WDR_OK;
    int rc;
    if ((rc = pw.reserve(M, d)) != WDR_OK) return rc;
    const int64_t Mc = pw.cap_rows;
    WDR_CUDA_TRY(cudaMemcpyAsync(pw.row_b, row_b.data(), sizeof(int32_t) * M, cudaMemcpyHostToDevice, st));
    WDR_CUDA_TRY(cudaMemcpyAsync(pw.row_pos, row_pos.data(), sizeof(int32_t) * M, cudaMemcpyHostToDevice, st));
    WDR_CUDA_TRY(cudaMemcpyAsync(pw.row_off, row_off.data(), sizeof(int32_t) * kDecMaxWindows, cudaMemcpyHostToDevice, st));
    WDR_CUDA_TRY(cudaStreamSynchronize(st));  // the staging vectors die with this frame
    static DeviceOnce attr_once;
    // shared arrays sized by this batch's longest sequence, not by the 256-token bound: at ~130 tokens three CTAs fit an SM instead of one
    const int smem_self_max = (int)sizeof(float) * (2 * kDtwpMaxT * 65 + 8 * 64 + 8 * kDtwpMaxT);
    int cap_T = 8;
    for (int b = 0; b < B; b++) cap_T = std::max(cap_T, (int)T_host[b]);
    cap_T = (cap_T + 7) / 8 * 8;
    const int smem_self = (int)sizeof(float) * (2 * cap_T * 65 + 8 * 64 + 8 * cap_T);
    const int smem_cross = (int)sizeof(float) * (kDtwpQB * 64 + kDtwpQB * kDtwpPStride);
    const int smem_cross_mma = (int)sizeof(float) * (kDtwpQB * kMqQPitch + kDtwpQB * kMqPitch);
    static const bool cross_ffma = getenv("WDR_DTWP_FFMA") != nullptr;  // A/B knob: the fp32-FMA kernel instead of the tensor-core one
    WDR_CUDA_TRY(per_device_once(attr_once, [&] {
        cudaError_t e = cudaFuncSetAttribute(dtwp_self_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_self_max);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(dtwp_cross_attn_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_cross_mma);
        return e != cudaSuccess ? e : cudaFuncSetAttribute(dtwp_cross_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_cross);
    }));
    int L_run = 0;
    for (auto& lh : ctx->aheads) L_run = std::max(L_run, lh.first + 1);
    L_run = std::min(L_run, a.n_dec_layer);
    {
        ProfScope ps(prof, KC_DECODER, st);
        dtwp_embed_kernel<<<(unsigned)M, 128, 0, st>>>(ws.seq, pw.row_b, pw.row_pos, w.tok_emb, w.dec_pos, d, a.n_vocab, pw.x);
        WDR_LAUNCH_CHECK();
    }
    const float* pend_bias = nullptr;
    bool pending = false;
    auto ln = [&](const float* g, const float* bta) -> int {
        ProfScope ps(prof, KC_DECODER, st);
        dec_ln_kernel<<<(unsigned)M, d / 4, 0, st>>>(pw.x, pending ? pw.part : nullptr, 1, 0, d, pend_bias, g, bta, pw.h, Mc * d, d);
        WDR_LAUNCH_CHECK();
        pending = false;
        return WDR_OK;
    };
    for (int l = 0; l < L_run; l++) {
        const DecLayerW& e = w.dec[l];
        if ((rc = ln(e.ln1_g, e.ln1_b)) != WDR_OK) return rc;
        if ((rc = packed_gemm(pw.h, M, Mc, e.w_qkv, 3 * d, d, pw.part, st, prof)) != WDR_OK) return rc;
        {
            ProfScope ps(prof, KC_DECODER, st);
            dtwp_self_attn_kernel<<<dim3(H, B), 256, smem_self, st>>>(pw.part, e.b_qkv, pw.row_off, ws.aw_T, d, pw.att, Mc * d, cap_T);
            WDR_LAUNCH_CHECK();
        }
        if ((rc = packed_gemm(pw.att, M, Mc, e.w_o, d, d, pw.part, st, prof)) != WDR_OK) return rc;
        pending = true; pend_bias = e.b_o;
        if ((rc = ln(e.ln2_g, e.ln2_b)) != WDR_OK) return rc;
        if ((rc = packed_gemm(pw.h, M, Mc, e.w_cq, d, d, pw.part, st, prof)) != WDR_OK) return rc;
        {
            ProfScope ps(prof, KC_DEC_CROSS_BATCHED, st);
            if (cross_ffma)
                dtwp_cross_attn_kernel<<<dim3((max_T + kDtwpQB - 1) / kDtwpQB, H, B), 256, smem_cross, st>>>(
                    pw.part, e.b_cq, ws.ckv[l], d, pw.row_off, ws.aw_T, pw.att, Mc * d, ws.ahead_map + (size_t)l * H, ws.aw, ws.aw_off, ws.aw_A);
            else
                dtwp_cross_attn_mma_kernel<<<dim3((max_T + kDtwpQB - 1) / kDtwpQB, H, B), 256, smem_cross_mma, st>>>(
                    pw.part, e.b_cq, ws.ckv[l], d, pw.row_off, ws.aw_T, pw.att, Mc * d, ws.ahead_map + (size_t)l * H, ws.aw, ws.aw_off, ws.aw_A);
            WDR_LAUNCH_CHECK();
        }
        if (l == L_run - 1) break;  // nothing after the last alignment head's probabilities is used
        if ((rc = packed_gemm(pw.att, M, Mc, e.w_co, d, d, pw.part, st, prof)) != WDR_OK) return rc;
        pending = true; pend_bias = e.b_co;
        if ((rc = ln(e.ln3_g, e.ln3_b)) != WDR_OK) return rc;
        {
            GemmDesc g;
            g.A = pw.h; g.a_row_stride = d; g.rows_per_batch = (int)M; g.n_batch = 1;
            g.W = e.w_fc1; g.ldw = d; g.N = 4 * d; g.K = d;
            g.epilogue = EPI_BIAS_GELU_SPLIT; g.out = pw.ff; g.ldc = 4 * d; g.bias = e.b_fc1; g.bn = 64;
            g.dual_a = true; g.a_dual_stride = Mc * d;
            g.split_stride = Mc * 4 * d;
            g.w_kb_major = true;
            ProfScope ps(prof, KC_DEC_GEMM, st);
            if ((rc = gemm_bf16(g, st)) != WDR_OK) return rc;
        }
        if ((rc = packed_gemm(pw.ff, M, Mc, e.w_fc2, d, 4 * d, pw.part, st, prof)) != WDR_OK) return rc;
        pending = true; pend_bias = e.b_fc2;
    }
    return WDR_OK;
}

int decoder_sample(const wdr_context* ctx, DecoderWorkspace& ws, int B, int pos, const SampleParams& sp, cudaStream_t st, Profiler* prof,
                   bool pos_on_device, bool pdl) {
    WDR_REQUIRE(sp.n_vocab <= kSampThreads * kSampPer, "vocabulary larger than the sampler's register tile");
    ProfScope ps(prof, KC_DECODER, st);
    WDR_CUDA_TRY(launch_kernel(dec_sample_kernel, dim3(B), dim3(kSampThreads), 0, st, pdl, ws.logits, ws.ldv, ws.win, ws.tokens, ws.seq,
                               pos_on_device ? ws.pos_dev : nullptr, pos, sp, ws.done_count));
    WDR_LAUNCH_CHECK();
    return WDR_OK;
}

int decoder_topk(const wdr_context* ctx, DecoderWorkspace& ws, int R, const SampleParams& sp, int k_top, float temperature, float* probs_out,
                 cudaStream_t st, Profiler* prof) {
    WDR_REQUIRE(sp.n_vocab <= kSampThreads * kSampPer && k_top >= 1 && k_top <= kBeamMax && R <= kDecMaxBatch, "decoder_topk: bad arguments");
    ProfScope ps(prof, KC_DECODER, st);
    dec_topk_kernel<<<R, kSampThreads, 0, st>>>(ws.logits, ws.ldv, ws.beam_rows, sp, k_top, temperature, ws.beam_cands, ws.beam_nosp, probs_out);
    WDR_LAUNCH_CHECK();
    return WDR_OK;
}

int decoder_beam_reorder(DecoderWorkspace& ws, int R, int pos_last, cudaStream_t st) {
    int32_t* old_t = ws.beam_anc_cur;
    int32_t* new_t = old_t == ws.beam_anc[0] ? ws.beam_anc[1] : ws.beam_anc[0];
    beam_anc_kernel<<<R, 128, 0, st>>>(old_t, new_t, ws.beam_parent, R, kDecSeqCap, pos_last);
    WDR_LAUNCH_CHECK();
    ws.beam_anc_cur = new_t;
    return WDR_OK;
}

// One greedy iteration as a replayable graph: sample at *pos_dev, advance the counter, decoder step (with logits) at the new
// position.  Captured once per (batch, sampling parameters) on the caller's stream and cached in the workspace.
int decoder_decode_graph(const wdr_context* ctx, DecoderWorkspace& ws, int B, const SampleParams& sp, cudaStream_t st, cudaGraphExec_t* out) {
    if (ws.step_graph && ws.graph_B == B && memcmp(&ws.graph_sp, &sp, sizeof(sp)) == 0) { *out = ws.step_graph; return WDR_OK; }
    if (ws.step_graph) { cudaGraphExecDestroy(ws.step_graph); ws.step_graph = nullptr; }
    WDR_CUDA_TRY(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    // WDR_PDL=1: every kernel of the graph but the first becomes a programmatic dependent launch — its CTAs are scheduled while
    // the predecessor drains, run their prologue (the GEMMs also prefetch their first weight tiles) and then wait for it.
    // Measured on B200 (large-v3, 120 windows, 220 iterations): 1981 ms with PDL vs 1955 ms without — inside a graph the
    // kernel-to-kernel gap is already ~0.4 us and the early CTAs only take SM resources from the draining kernel — so it is off.
    static const bool pdl = getenv("WDR_PDL") != nullptr;
    int rc = decoder_sample(ctx, ws, B, 0, sp, st, nullptr, true, false);
    if (rc == WDR_OK) {
        const cudaError_t le = launch_kernel(dec_advance_kernel, dim3(1), dim3(1), 0, st, pdl, ws.pos_dev);
        count_launch();
        rc = le == cudaSuccess ? decoder_step(ctx, ws, B, 0, true, DEC_MODE_DECODE, st, nullptr, true, pdl) : WDR_ERR_CUDA;
    }
    cudaGraph_t graph = nullptr;
    const cudaError_t e = cudaStreamEndCapture(st, &graph);
    if (rc != WDR_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess || !graph) { set_error("decode graph capture failed: %s", cudaGetErrorString(e)); cudaGetLastError(); return WDR_ERR_CUDA; }
    size_t n_nodes = 0;
    cudaGraphGetNodes(graph, nullptr, &n_nodes);
    const cudaError_t e2 = cudaGraphInstantiate(&ws.step_graph, graph, 0);
    cudaGraphDestroy(graph);
    if (e2 != cudaSuccess) { ws.step_graph = nullptr; set_error("cudaGraphInstantiate: %s", cudaGetErrorString(e2)); return WDR_ERR_CUDA; }
    ws.graph_B = B;
    ws.graph_sp = sp;
    ws.graph_nodes = (int)n_nodes;
    *out = ws.step_graph;
    return WDR_OK;
}

}  // namespace wdr
