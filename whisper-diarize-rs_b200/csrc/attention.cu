// attention.cu — encoder self-attention (non-causal, T = 1500, d_head = 64) on tcgen05 tensor cores.
//
// Replaces the QK^T -> soft_max -> V matmul chain of whisper.cpp's encoder graph (whisper_encode_internal,
// SURVEY A.2; reached from reference src/transcribe.rs:389).
//
// One CTA = one (chunk, head, 128-query tile); 6 warps: warp 0 = TMA producer, warp 1 = MMA issuer + TMEM
// owner, warps 2..5 = softmax (one thread per query row == one TMEM lane).  Per 128-key tile:
//   S = Q K^T  (tcgen05.mma 128x128x16 x4, accumulator in TMEM),
//   ONE read of the S row into registers (4 x tcgen05.ld.x32 in flight together); the S buffer is released at once, so the MMA
//   warp computes S of the next tile underneath this tile's softmax,
//   row max, p = exp2(s * scale*log2e - m_ref), row sum, bf16 P written to shared memory in the 128-byte-swizzled K-major
//   layout the MMA reads,
//   O += P V  (tcgen05.mma 128x64x16 x8; V is consumed K-major from the V^T buffer the QKV GEMM epilogue wrote): the output
//   accumulator stays in TMEM across the key tiles.
// m_ref is the reference maximum of the row.  It follows the running maximum lazily (FlashAttention-4): while a tile's maximum
// exceeds it by no more than 8 in the exp2 domain (p <= 256, harmless in bf16 / fp32) nothing is rescaled; otherwise the row's O
// is read from TMEM, scaled and written back (tcgen05.st) and the row sum is scaled with it — after the first tiles that is rare.
// The result is the exact softmax of the row: numerator and denominator use the same m_ref.
// Two CTAs are resident per SM (97 KB shared memory, 256 TMEM columns each) so one CTA's tensor work overlaps
// the other's softmax.  Keys beyond T in the last tile are masked to -inf; query rows beyond T are not stored.
#include "common.cuh"
#include "sm100.cuh"

namespace wdr {

using namespace sm100;

int make_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box);

constexpr int kAttThreads = 192;
constexpr int kAttQ = 128, kAttK = 128, kAttD = 64;
constexpr int kQBytes = kAttQ * kAttD * 2;      // 16 KB
constexpr int kKBytes = kAttK * kAttD * 2;      // 16 KB
constexpr int kVBytes = kAttD * kAttK * 2;      // 16 KB (two 8 KB K-blocks of [64 d][64 keys])
constexpr int kPBytes = kAttQ * kAttK * 2;      // 32 KB (two 16 KB K-blocks of [128 q][64 keys])
constexpr int kAttSmem = kQBytes + 2 * kKBytes + kVBytes + kPBytes + 1024;
constexpr uint32_t kAttTmemCols = 256;

struct AttParams {
    int T;          // tokens per chunk (1500)
    int T_pad;      // per-chunk column stride of V^T (multiple of 8: TMA box starts must be 16-byte aligned)
    int n_kt;       // key tiles per chunk
    int d_model;
    __nv_bfloat16* out;  // [B*T][d_model]
    float scale_log2e;   // (1/sqrt(64)) * log2(e)
};

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

constexpr float kAttRescaleThreshold = 8.0f;  // exp2 domain: p stays <= 2^8 against a stale reference maximum

// row max of one 32-column chunk held in registers (MASK = tile has padding keys)
template <bool MASK>
__device__ __forceinline__ float att_chunk_max(const uint32_t (&r)[32], int c, int n_valid, float m) {
    float ma = m, mb = -INFINITY;
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
        float x0 = __uint_as_float(r[i]), x1 = __uint_as_float(r[i + 1]), x2 = __uint_as_float(r[i + 2]), x3 = __uint_as_float(r[i + 3]);
        if (MASK) {
            if (c * 32 + i >= n_valid) x0 = -INFINITY;
            if (c * 32 + i + 1 >= n_valid) x1 = -INFINITY;
            if (c * 32 + i + 2 >= n_valid) x2 = -INFINITY;
            if (c * 32 + i + 3 >= n_valid) x3 = -INFINITY;
        }
        ma = max3(ma, x0, x1);
        mb = max3(mb, x2, x3);
    }
    return fmaxf(ma, mb);
}

// p = 2^(s*scale - m_scaled) of one chunk -> bf16 P in shared memory (128-byte-swizzled K-major rows); adds to the row sum
template <bool MASK>
__device__ __forceinline__ void att_chunk_probs(const uint32_t (&r)[32], int c, int n_valid, float scale, float m_scaled, unsigned char* p_row, int sw,
                                                float& l0, float& l1) {
    uint32_t packed[16];
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
        float p0 = ex2_approx(fmaf(__uint_as_float(r[i]), scale, -m_scaled));
        float p1 = ex2_approx(fmaf(__uint_as_float(r[i + 1]), scale, -m_scaled));
        if (MASK) {
            if (c * 32 + i >= n_valid) p0 = 0.0f;
            if (c * 32 + i + 1 >= n_valid) p1 = 0.0f;
        }
        l0 += p0;
        l1 += p1;
        __nv_bfloat162 pk = __floats2bfloat162_rn(p0, p1);
        packed[i >> 1] = *reinterpret_cast<uint32_t*>(&pk);
    }
    unsigned char* blk = p_row + (c >> 1) * (kPBytes / 2);
#pragma unroll
    for (int g = 0; g < 4; g++) {
        const int chunk = (c & 1) * 4 + g;  // 16-byte chunk within the 128-byte row
        *reinterpret_cast<uint4*>(blk + ((chunk ^ sw) << 4)) = make_uint4(packed[4 * g], packed[4 * g + 1], packed[4 * g + 2], packed[4 * g + 3]);
    }
}

__global__ void __launch_bounds__(kAttThreads, 2)
encoder_attention_kernel(const __grid_constant__ CUtensorMap tma_qk, const __grid_constant__ CUtensorMap tma_vt, const AttParams p) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    unsigned char* sQ = smem;
    unsigned char* sK = sQ + kQBytes;
    unsigned char* sV = sK + 2 * kKBytes;
    unsigned char* sP = sV + kVBytes;
    __shared__ __align__(8) uint64_t bar_q, bar_kfull[2], bar_kempty[2], bar_vfull, bar_vempty, bar_s, bar_sfree, bar_p, bar_o;
    __shared__ uint32_t s_tmem_base;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tma_qk);
        tma_prefetch_desc(&tma_vt);
        mbar_init(&bar_q, 1);
        for (int s = 0; s < 2; s++) { mbar_init(&bar_kfull[s], 1); mbar_init(&bar_kempty[s], 1); }
        mbar_init(&bar_vfull, 1);
        mbar_init(&bar_vempty, 1);
        mbar_init(&bar_s, 1);
        mbar_init(&bar_sfree, 128);
        mbar_init(&bar_p, 128);
        mbar_init(&bar_o, 1);
        mbar_fence_init();
    }
    if (warp == 1) {
        tmem_alloc(&s_tmem_base, kAttTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem_base;
    const uint32_t tmem_s = tmem_base;          // 128 columns
    const uint32_t tmem_o = tmem_base + 128;    // 64 columns

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(&bar_q, kQBytes);
            tma_load_3d(sQ, &tma_qk, &bar_q, h * kAttD, qt * kAttQ, b);
            for (int j = 0; j < p.n_kt; j++) {
                const int s = j & 1;
                mbar_wait_parked(&bar_kempty[s], ((j >> 1) & 1) ^ 1);
                mbar_arrive_expect_tx(&bar_kfull[s], kKBytes);
                tma_load_3d(sK + s * kKBytes, &tma_qk, &bar_kfull[s], p.d_model + h * kAttD, j * kAttK, b);
                mbar_wait_parked(&bar_vempty, (j & 1) ^ 1);
                mbar_arrive_expect_tx(&bar_vfull, kVBytes);
                const int tok0 = b * p.T_pad + j * kAttK;
                tma_load_2d(sV, &tma_vt, &bar_vfull, tok0, h * kAttD);
                tma_load_2d(sV + kVBytes / 2, &tma_vt, &bar_vfull, tok0 + 64, h * kAttD);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc_s = umma_idesc_bf16(kAttQ, kAttK);
            constexpr uint32_t idesc_o = umma_idesc_bf16(kAttQ, kAttD);
            mbar_wait_parked(&bar_q, 0);
            const uint64_t dq = umma_desc_kmajor_sw128(smem_u32(sQ));
            const uint64_t dp0 = umma_desc_kmajor_sw128(smem_u32(sP));
            const uint64_t dp1 = umma_desc_kmajor_sw128(smem_u32(sP + kPBytes / 2));
            const uint64_t dv0 = umma_desc_kmajor_sw128(smem_u32(sV));
            const uint64_t dv1 = umma_desc_kmajor_sw128(smem_u32(sV + kVBytes / 2));
            auto issue_s = [&](int j) {  // S(j) = Q K(j)^T into the one S buffer
                const int s = j & 1;
                mbar_wait_parked(&bar_kfull[s], (j >> 1) & 1);
                tc_fence_after();
                const uint64_t dk = umma_desc_kmajor_sw128(smem_u32(sK + s * kKBytes));
#pragma unroll
                for (int k = 0; k < kAttD / 16; k++) umma_bf16(tmem_s, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
                umma_commit(&bar_s);
                umma_commit(&bar_kempty[s]);
            };
            issue_s(0);
            for (int j = 0; j < p.n_kt; j++) {
                // the softmax threads hold S(j) in registers: the next S goes underneath their arithmetic
                mbar_wait_parked(&bar_sfree, j & 1);
                tc_fence_after();
                if (j + 1 < p.n_kt) issue_s(j + 1);
                mbar_wait_parked(&bar_p, j & 1);
                mbar_wait_parked(&bar_vfull, j & 1);
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < 4; k++) umma_bf16(tmem_o, dp0 + 2 * k, dv0 + 2 * k, idesc_o, (j != 0) | (k != 0));  // O stays in TMEM
#pragma unroll
                for (int k = 0; k < 4; k++) umma_bf16(tmem_o, dp1 + 2 * k, dv1 + 2 * k, idesc_o, 1);
                umma_commit(&bar_o);
                umma_commit(&bar_vempty);
            }
        }
    } else {
        // ===================== softmax / output: thread == query row == TMEM lane =====================
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        float m_ref = -INFINITY;  // reference maximum of the row, exp2 domain (score * scale_log2e)
        float l_run = 0.0f;
        unsigned char* p_row = sP + row * 128;
        const int sw = row & 7;
        for (int j = 0; j < p.n_kt; j++) {
            const int n_valid = min(kAttK, p.T - j * kAttK);  // keys of this tile that exist
            const bool full = (n_valid == kAttK);
            mbar_wait(&bar_s, j & 1);
            tc_fence_after();
            // the whole S row into registers, four loads in flight; then the S buffer belongs to the next tile
            uint32_t s0[32], s1[32], s2[32], s3[32];
            tmem_ld_32x32(tmem_s + lane_addr, s0);
            tmem_ld_32x32(tmem_s + lane_addr + 32, s1);
            tmem_ld_32x32(tmem_s + lane_addr + 64, s2);
            tmem_ld_32x32(tmem_s + lane_addr + 96, s3);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(&bar_sfree);
            float m0, m1, m2, m3;  // four independent max chains (one 64-long chain is pure dependency latency)
            if (full) {
                m0 = att_chunk_max<false>(s0, 0, n_valid, -INFINITY); m1 = att_chunk_max<false>(s1, 1, n_valid, -INFINITY);
                m2 = att_chunk_max<false>(s2, 2, n_valid, -INFINITY); m3 = att_chunk_max<false>(s3, 3, n_valid, -INFINITY);
            } else {
                m0 = att_chunk_max<true>(s0, 0, n_valid, -INFINITY); m1 = att_chunk_max<true>(s1, 1, n_valid, -INFINITY);
                m2 = att_chunk_max<true>(s2, 2, n_valid, -INFINITY); m3 = att_chunk_max<true>(s3, 3, n_valid, -INFINITY);
            }
            const float m_tile = fmaxf(max3(m0, m1, m2), m3);
            const float m_tile_scaled = m_tile * p.scale_log2e;
            // previous tile's P V must be finished before sP is overwritten (and before its O is touched)
            if (j > 0) {
                mbar_wait(&bar_o, (j - 1) & 1);
                tc_fence_after();
            }
            // lazy rescale: only when this tile's maximum leaves the window of the reference maximum
            const bool need = m_tile_scaled > m_ref + kAttRescaleThreshold;  // always true on the first tile (m_ref = -inf)
            if (j == 0) {
                m_ref = m_tile_scaled;
            } else if (__any_sync(0xffffffffu, need)) {  // tcgen05.ld / st are warp-collective: rows that need nothing scale by 1
                const float alpha = need ? ex2_approx(m_ref - m_tile_scaled) : 1.0f;
#pragma unroll 1
                for (int c = 0; c < kAttD / 8; c++) {  // 8 columns at a time: the S row (128 registers) is live here
                    uint32_t r[8];
                    tmem_ld_32x8(tmem_o + lane_addr + c * 8, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 8; i++) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
                    tmem_st_32x8(tmem_o + lane_addr + c * 8, r);
                }
                tmem_st_wait();
                l_run *= alpha;
                if (need) m_ref = m_tile_scaled;
            }
            // probabilities against m_ref -> bf16 P in shared memory (swizzled K-major), row sum
            float l0 = 0.0f, l1 = 0.0f;
            if (full) {
                att_chunk_probs<false>(s0, 0, n_valid, p.scale_log2e, m_ref, p_row, sw, l0, l1);
                att_chunk_probs<false>(s1, 1, n_valid, p.scale_log2e, m_ref, p_row, sw, l0, l1);
                att_chunk_probs<false>(s2, 2, n_valid, p.scale_log2e, m_ref, p_row, sw, l0, l1);
                att_chunk_probs<false>(s3, 3, n_valid, p.scale_log2e, m_ref, p_row, sw, l0, l1);
            } else {
                att_chunk_probs<true>(s0, 0, n_valid, p.scale_log2e, m_ref, p_row, sw, l0, l1);
                att_chunk_probs<true>(s1, 1, n_valid, p.scale_log2e, m_ref, p_row, sw, l0, l1);
                att_chunk_probs<true>(s2, 2, n_valid, p.scale_log2e, m_ref, p_row, sw, l0, l1);
                att_chunk_probs<true>(s3, 3, n_valid, p.scale_log2e, m_ref, p_row, sw, l0, l1);
            }
            l_run += l0 + l1;
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(&bar_p);
        }
        // the finished accumulator: O / l
        mbar_wait(&bar_o, (p.n_kt - 1) & 1);
        tc_fence_after();
        const int tok = qt * kAttQ + row;
        const float inv = 1.0f / l_run;
        __nv_bfloat16* o = p.out + ((int64_t)b * p.T + tok) * p.d_model + h * kAttD;
#pragma unroll
        for (int c = 0; c < kAttD / 32; c++) {
            uint32_t r[32];
            tmem_ld_32x32(tmem_o + lane_addr + c * 32, r);
            tmem_ld_wait();
            if (tok < p.T) {
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
                    uint4 w;
                    __nv_bfloat162 t0 = __floats2bfloat162_rn(__uint_as_float(r[i]) * inv, __uint_as_float(r[i + 1]) * inv);
                    __nv_bfloat162 t1 = __floats2bfloat162_rn(__uint_as_float(r[i + 2]) * inv, __uint_as_float(r[i + 3]) * inv);
                    __nv_bfloat162 t2 = __floats2bfloat162_rn(__uint_as_float(r[i + 4]) * inv, __uint_as_float(r[i + 5]) * inv);
                    __nv_bfloat162 t3 = __floats2bfloat162_rn(__uint_as_float(r[i + 6]) * inv, __uint_as_float(r[i + 7]) * inv);
                    w.x = *reinterpret_cast<uint32_t*>(&t0);
                    w.y = *reinterpret_cast<uint32_t*>(&t1);
                    w.z = *reinterpret_cast<uint32_t*>(&t2);
                    w.w = *reinterpret_cast<uint32_t*>(&t3);
                    *reinterpret_cast<uint4*>(o + c * 32 + i) = w;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kAttTmemCols);
    }
}

// qk: bf16 [B*T][2*d_model] (q | k); vt: bf16 [d_model][ldt], chunk b's tokens at columns b*T_pad .. b*T_pad+T-1 with
// T_pad = round_up(T, 8) (pad columns zero); out: bf16 [B*T][d_model]
int encoder_attention(const __nv_bfloat16* qk, const __nv_bfloat16* vt, int64_t ldt, int B, int T, int n_head, int d_model,
                      __nv_bfloat16* out, cudaStream_t st) {
    WDR_REQUIRE(d_model == n_head * kAttD, "d_head must be 64");
    const int T_pad = (T + 7) / 8 * 8;
    WDR_REQUIRE(ldt % 8 == 0 && ldt >= (int64_t)B * T_pad, "ldt must be a multiple of 8 covering B * round_up(T, 8) columns");
    CUtensorMap tqk, tvt;
    {
        const uint64_t dims[3] = {(uint64_t)2 * d_model, (uint64_t)T, (uint64_t)B};
        const uint64_t str[2] = {(uint64_t)2 * d_model * 2, (uint64_t)T * 2 * d_model * 2};
        const uint32_t box[3] = {kAttD, kAttQ, 1};
        int rc = make_tmap_bf16(&tqk, qk, 3, dims, str, box);
        if (rc != WDR_OK) return rc;
    }
    {
        const uint64_t dims[2] = {(uint64_t)B * T_pad, (uint64_t)d_model};
        const uint64_t str[1] = {(uint64_t)ldt * 2};
        const uint32_t box[2] = {64, kAttD};
        int rc = make_tmap_bf16(&tvt, vt, 2, dims, str, box);
        if (rc != WDR_OK) return rc;
    }
    static DeviceOnce attr_once;
    WDR_CUDA_TRY(per_device_once(attr_once, [] { return cudaFuncSetAttribute(encoder_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttSmem); }));
    AttParams p;
    p.T = T;
    p.T_pad = T_pad;
    p.n_kt = (T + kAttK - 1) / kAttK;
    p.d_model = d_model;
    p.out = out;
    p.scale_log2e = 0.125f * 1.4426950408889634f;
    dim3 grid((T + kAttQ - 1) / kAttQ, n_head, B);
    encoder_attention_kernel<<<grid, kAttThreads, kAttSmem, st>>>(tqk, tvt, p);
    WDR_LAUNCH_CHECK();
    return WDR_OK;
}

}  // namespace wdr

extern "C" int wdr_encoder_attention_dev(const uint16_t* qk, const uint16_t* vt, int64_t ldt, int n_chunks, int T, int n_head,
                                         int d_model, uint16_t* out, void* stream) {
    wdr::clear_error();
    int rc = wdr::ensure_device(-1);
    if (rc != WDR_OK) return rc;
    return wdr::encoder_attention(reinterpret_cast<const __nv_bfloat16*>(qk), reinterpret_cast<const __nv_bfloat16*>(vt), ldt, n_chunks,
                                  T, n_head, d_model, reinterpret_cast<__nv_bfloat16*>(out), (cudaStream_t)stream);
}
