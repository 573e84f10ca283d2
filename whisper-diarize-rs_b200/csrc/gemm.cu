// gemm.cu — persistent, warp-specialised bf16 GEMM on tcgen05 tensor cores for sm_100a.
//
// The Whisper encoder/decoder linear layers, the conv stem (as implicit GEMM over strided row views) and
// the WeSpeaker/pyannote dense layers all run through this kernel.  It replaces ggml's mul_mat (whisper.cpp
// `whisper_encode_internal`, reached from reference src/transcribe.rs:389) and ONNX Runtime's Gemm/Conv.
//
//   D[M x N] = A[M x K] * W[N x K]^T, bf16 operands (both K-major), fp32 accumulation in TMEM.
//
// CTA = 8 warps: warp 0 = TMA producer (cp.async.bulk.tensor, 128-byte swizzle, kStages-deep mbarrier ring),
// warp 1 = MMA issuer (one elected thread, tcgen05.mma cta_group::1 128 x BN x 16, tcgen05.commit frees the
// shared-memory stage / publishes the accumulator), warp 2 = TMEM allocator, warps 4..7 = epilogue
// (tcgen05.ld 32 lanes x 32 columns per warp, fused bias / GELU / residual / positional embedding /
// transposed store).  Two accumulator stages in TMEM let the epilogue of tile i overlap the MMAs of tile i+1.
// One CTA per SM, persistent over a (m-tile, n-tile) list with n fastest so that concurrently running CTAs
// share the same A tile through L2.
#include <cuda.h>
#include <mutex>
#include "common.cuh"
#include "gemm.cuh"
#include "sm100.cuh"

namespace wdr {

using namespace sm100;

constexpr int kBM = 128;
constexpr int kBK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int kGemmThreads = 384;  // 4 control warps + 8 epilogue warps

struct GemmParams {
    int rows_per_batch, n_batch, N, K;
    int kb_per_tap;
    int tiles_per_batch, n_tiles, total_tiles;
    int splits, kb_per_split;  // split-K: work item t -> (tile t / splits, split t % splits); split s covers k-blocks [s*kb_per_split, ...)
    int64_t split_stride;      // elements between the partial outputs of consecutive splits
    void* out;
    int64_t ldc;
    int64_t c_batch_stride;
    const float* bias;
    const float* resid;
    const __nv_bfloat16* resid_bf16;
    const float* pos;
    __nv_bfloat16* out_t;
    int64_t ldt;
    int64_t t_batch_stride;
    int n_split;
    int group_rows;
    int w_kb_major;  // B operand coordinates: (0, kb*N + n) instead of (kb*64, n)
    const int32_t* tile_pitch;  // CONV: per M tile, pitch of the padded map that owns it / row where that map's block starts
    const int32_t* tile_row0;
    int conv_F;
    int dbg_no_a;    // measurement aid (WDR_DEBUG_GEMM_NO_A, dual-A GEMMs only): the activation tiles are not loaded at all — WRONG results;
                     // what remains is the weight stream alone, i.e. the most any activation-traffic optimisation could win
};

__device__ __forceinline__ float gelu_tanh(float x) {
    // 0.5 x (1 + tanh(u)),  u = sqrt(2/pi) * x * (1 + 0.044715 x^2)   (ggml_gelu_f32); one MUFU (tanh.approx, abs err 2^-11)
    const float u = 0.7978845608028654f * x * fmaf(0.044715f * x, x, 1.0f);
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
    const float hx = 0.5f * x;
    return fmaf(hx, t, hx);
}

__device__ __forceinline__ float gelu_tanh_precise(float x) {
    return 0.5f * x * (1.0f + tanhf(0.79788456080286535587989211986876f * x * (1.0f + 0.044715f * x * x)));
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}

// CONV (gemm.cuh, conv2d): 1 / 2 = the A rows of k-block kb are the tile's rows shifted by (dy * pitch + dx) — the 3 x 3 convolution as an
// implicit GEMM over a zero-padded map; 1, 2, 3 = the epilogue zeroes the non-interior rows of the output map.
// A row shift of k-block kb (and its column block) for the implicit convolution
template <int CONV>
__device__ __forceinline__ void conv_a_coord(int kb, int kb_per_tap, int pitch, int& kcol, int& shift) {
    if (CONV == 1) {
        const int tap = kb / kb_per_tap;
        kcol = kb - tap * kb_per_tap;
        const int ky = tap / 3;
        shift = (ky - 1) * pitch + (tap - ky * 3) - 1;
    } else {  // CONV == 2: (dy, pair of taps)
        kcol = 0;
        shift = ((kb >> 1) - 1) * pitch + ((kb & 1) << 1) - 1;
    }
}

template <int BN, int STAGES, int EPI, bool DUAL, int CONV = 0>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, const GemmParams p) {
    constexpr int kABytes = kBM * kBK * 2;
    constexpr int kBBytes = BN * kBK * 2;
    constexpr uint32_t kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
    static_assert(2 * BN <= 512, "two accumulator stages must fit TMEM");
    static_assert(BN % 32 == 0 && BN >= 32 && BN <= 256, "BN");

    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    // DUAL: the A operand is a (hi, lo) bf16 pair (A = hi + lo to 16 mantissa bits); both halves multiply the same W tile
    constexpr int kAStage = DUAL ? 2 * kABytes : kABytes;
    unsigned char* sA = smem;
    unsigned char* sB = smem + STAGES * kAStage;
    __shared__ __align__(8) uint64_t bar_full[STAGES];
    __shared__ __align__(8) uint64_t bar_empty[STAGES];
    __shared__ __align__(8) uint64_t bar_tfull[2];
    __shared__ __align__(8) uint64_t bar_tempty[2];
    __shared__ uint32_t s_tmem_base;
    __shared__ __align__(16) float4 s_stage[8][32 * 8];  // per epilogue warp: 32 rows x 32 fp32

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_kb = (p.K + kBK - 1) / kBK;
    pdl_launch_dependents();

    // The producer thread initialises the operand barriers itself and requests the first item's first STAGES k-blocks at once —
    // before the TMEM allocation and the block-wide barrier of the prologue, so their latency (tensor-map fetch + HBM / L2) runs
    // under it; the decode GEMMs are a few microseconds long and one such round trip is a visible share of them.
    int prefetched = 0;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tma_a);
        tma_prefetch_desc(&tma_b);
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_empty[s], 1);
        }
        mbar_fence_init();
        if (blockIdx.x < p.total_tiles) {
            const int w = blockIdx.x;
            const int t = w / p.splits, sp = w - t * p.splits;
            const int m_tile = t / p.n_tiles, n_tile = t - m_tile * p.n_tiles;
            const int batch = m_tile / p.tiles_per_batch, mt = m_tile - batch * p.tiles_per_batch;
            const int kb0 = sp * p.kb_per_split, kb1 = min(num_kb, kb0 + p.kb_per_split);
            prefetched = min(STAGES, kb1 - kb0);
            // The weight tiles never depend on the predecessor kernel: they are requested (from HBM: the long latency) BEFORE the
            // programmatic-dependency wait, the activation tiles (L2-resident, just written by the predecessor) after it.
            for (int i = 0; i < prefetched; i++) {
                const int kb = kb0 + i;
                mbar_arrive_expect_tx(&bar_full[i], (DUAL && p.dbg_no_a) ? kBBytes : kAStage + kBBytes);
                if (p.w_kb_major) tma_load_2d(sB + i * kBBytes, &tma_b, &bar_full[i], 0, kb * p.N + n_tile * BN);
                else tma_load_2d(sB + i * kBBytes, &tma_b, &bar_full[i], kb * kBK, n_tile * BN);
            }
            pdl_wait();  // (a no-op unless launched as a programmatic dependent: then A must wait for the predecessor)
            const int pitch0 = (CONV == 1 || CONV == 2) ? p.tile_pitch[m_tile] : 0;
            for (int i = 0; i < prefetched && !(DUAL && p.dbg_no_a); i++) {
                const int kb = kb0 + i;
                int tap = kb / p.kb_per_tap, kcol = kb - tap * p.kb_per_tap;
                if (CONV == 1 || CONV == 2) conv_a_coord<CONV>(kb, p.kb_per_tap, pitch0, kcol, tap);
                tma_load_3d(sA + i * kAStage, &tma_a, &bar_full[i], kcol * kBK, mt * kBM + tap, batch);
                if (DUAL) tma_load_3d(sA + i * kAStage + kABytes, &tma_a, &bar_full[i], kcol * kBK, mt * kBM + tap, 1);
            }
        }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < 2; s++) {
            mbar_init(&bar_tfull[s], 1);
            mbar_init(&bar_tempty[s], 8);
        }
        mbar_fence_init();
    }
    if (warp == 2) {
        tmem_alloc(&s_tmem_base, kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem_base;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            // CONV: the pitch of a tile's map comes from global memory; it is fetched one tile ahead so that its latency never sits
            // between two tiles' loads
            int pitch_next = ((CONV == 1 || CONV == 2) && blockIdx.x < p.total_tiles) ? p.tile_pitch[(blockIdx.x / p.splits) / p.n_tiles] : 0;
            for (int w = blockIdx.x; w < p.total_tiles; w += gridDim.x) {
                const int t = w / p.splits, sp = w - t * p.splits;
                const int m_tile = t / p.n_tiles, n_tile = t - m_tile * p.n_tiles;
                const int batch = m_tile / p.tiles_per_batch, mt = m_tile - batch * p.tiles_per_batch;
                const int kb0 = sp * p.kb_per_split, kb1 = min(num_kb, kb0 + p.kb_per_split);
                const int pitch = pitch_next;
                if ((CONV == 1 || CONV == 2) && w + (int)gridDim.x < p.total_tiles) pitch_next = p.tile_pitch[((w + (int)gridDim.x) / p.splits) / p.n_tiles];
                for (int kb = kb0; kb < kb1; kb++) {
                    if (prefetched > 0) prefetched--;  // requested in the prologue: this stage is armed and both tiles are in flight
                    else {
                        mbar_wait(&bar_empty[s], ph ^ 1);
                        mbar_arrive_expect_tx(&bar_full[s], (DUAL && p.dbg_no_a) ? kBBytes : kAStage + kBBytes);
                        int tap = kb / p.kb_per_tap, kcol = kb - tap * p.kb_per_tap;
                        if (CONV == 1 || CONV == 2) conv_a_coord<CONV>(kb, p.kb_per_tap, pitch, kcol, tap);
                        if (!(DUAL && p.dbg_no_a)) tma_load_3d(sA + s * kAStage, &tma_a, &bar_full[s], kcol * kBK, mt * kBM + tap, batch);
                        if (DUAL && !p.dbg_no_a) tma_load_3d(sA + s * kAStage + kABytes, &tma_a, &bar_full[s], kcol * kBK, mt * kBM + tap, 1);
                        if (p.w_kb_major) tma_load_2d(sB + s * kBBytes, &tma_b, &bar_full[s], 0, kb * p.N + n_tile * BN);
                        else tma_load_2d(sB + s * kBBytes, &tma_b, &bar_full[s], kb * kBK, n_tile * BN);
                    }
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(kBM, BN);
            int s = 0;
            uint32_t ph = 0;
            int as = 0;
            uint32_t aph = 0;
            for (int w = blockIdx.x; w < p.total_tiles; w += gridDim.x) {
                const int sp = w % p.splits;
                const int kb0 = sp * p.kb_per_split, kb1 = min(num_kb, kb0 + p.kb_per_split);
                mbar_wait(&bar_tempty[as], aph ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * BN;
                for (int kb = kb0; kb < kb1; kb++) {
                    mbar_wait(&bar_full[s], ph);
                    tc_fence_after();
                    const uint64_t da = umma_desc_kmajor_sw128(smem_u32(sA + s * kAStage));
                    const uint64_t db = umma_desc_kmajor_sw128(smem_u32(sB + s * kBBytes));
#pragma unroll
                    for (int k = 0; k < kBK / 16; k++) {
                        // advance 16 bf16 = 32 B along K inside the 128 B swizzle row: +2 in the (addr >> 4) field
                        umma_bf16(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb != kb0) || (k != 0));
                    }
                    if (DUAL) {
                        const uint64_t dl = umma_desc_kmajor_sw128(smem_u32(sA + s * kAStage + kABytes));
#pragma unroll
                        for (int k = 0; k < kBK / 16; k++) umma_bf16(tmem_d, dl + 2 * k, db + 2 * k, idesc, 1);
                    }
                    umma_commit(&bar_empty[s]);
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
                umma_commit(&bar_tfull[as]);
                as ^= 1;
                if (as == 0) aph ^= 1;
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue: 8 warps = 4 TMEM lane quarters x 2 column halves =====================
        pdl_wait();  // the epilogue reads residuals and overwrites buffers the predecessor may still be reading
        const int q = warp & 3;           // TMEM lane group this warp may access
        const int half = (warp - 4) >> 2;  // which half of the tile's columns
        float4* stage = s_stage[warp - 4];
        const int g = lane & 7, rsub = lane >> 3;
        int as = 0;
        uint32_t aph = 0;
        for (int w = blockIdx.x; w < p.total_tiles; w += gridDim.x) {
            const int t = w / p.splits, sp = w - t * p.splits;
            const int m_tile = t / p.n_tiles, n_tile = t - m_tile * p.n_tiles;
            const int batch = m_tile / p.tiles_per_batch, mt = m_tile - batch * p.tiles_per_batch;
            const int r_in_batch = mt * kBM + q * 32 + lane;
            const bool row_ok = r_in_batch < p.rows_per_batch;
            const int64_t c_base = (int64_t)sp * p.split_stride + (int64_t)batch * p.c_batch_stride + (int64_t)(mt * kBM + q * 32) * p.ldc;
            uint32_t interior = 0xffu;  // CONV: bit `it` = row it * 4 + rsub of this warp's 32 is an interior position of its padded map
            if (CONV != 0) {
                const int P = p.tile_pitch[m_tile], local0 = mt * kBM + q * 32 - p.tile_row0[m_tile];
                interior = 0u;
#pragma unroll
                for (int it = 0; it < 8; it++) {
                    const int local = local0 + it * 4 + rsub, f1 = local / P, tt = local - f1 * P;
                    if (f1 >= 1 && f1 <= p.conv_F && tt < P - 1) interior |= 1u << it;
                }
            }
            bool waited = false;
#pragma unroll 1
            for (int c = half * (BN / 64); c < (half + 1) * (BN / 64); c++) {
                const int n0 = n_tile * BN + c * 32;
                if (n0 >= p.N) break;  // warp-uniform
                const int n = n0 + 4 * g;
                const bool col_ok = n < p.N;
                // operands that do not depend on the accumulator are fetched before waiting for it
                float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (p.bias && col_ok) bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + n));
                uint2 rres[8];  // EPI_BIAS_ADD_RELU_BF16: the residual's four bf16 of every row this lane stores
                if (EPI == EPI_BIAS_ADD_RELU_BF16) {
#pragma unroll
                    for (int it = 0; it < 8; it++) {
                        const int row = it * 4 + rsub;
                        rres[it] = make_uint2(0u, 0u);
                        if (mt * kBM + q * 32 + row < p.rows_per_batch && col_ok) rres[it] = *reinterpret_cast<const uint2*>(p.resid_bf16 + c_base + (int64_t)row * p.ldc + n);
                    }
                }
                float4 extra[8];
                if (EPI == EPI_BIAS_RESID_F32 || EPI == EPI_BIAS_GELU_POS_F32) {
#pragma unroll
                    for (int it = 0; it < 8; it++) {
                        const int row = it * 4 + rsub;
                        const int rib = mt * kBM + q * 32 + row;
                        extra[it] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (rib < p.rows_per_batch && col_ok) {
                            if (EPI == EPI_BIAS_RESID_F32) extra[it] = *reinterpret_cast<const float4*>(p.resid + c_base + (int64_t)row * p.ldc + n);
                            else extra[it] = __ldg(reinterpret_cast<const float4*>(p.pos + (int64_t)rib * p.N + n));
                        }
                    }
                }
                if (!waited) {
                    mbar_wait(&bar_tfull[as], aph);
                    tc_fence_after();
                    waited = true;
                }
                uint32_t r[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN + c * 32), r);
                tmem_ld_wait();
                if (EPI == EPI_QKV_BF16 && n0 >= p.n_split) {
                    // transposed store (V^T): lanes hold consecutive rows -> 64 B coalesced per column, straight from registers
                    if (row_ok) {
                        __nv_bfloat16* o = p.out_t + (int64_t)(n0 - p.n_split) * p.ldt + (int64_t)batch * p.t_batch_stride + r_in_batch;
#pragma unroll
                        for (int j = 0; j < 32; j++)
                            if (n0 + j < p.N) o[(int64_t)j * p.ldt] = __float2bfloat16_rn(__uint_as_float(r[j]) + (p.bias ? __ldg(p.bias + n0 + j) : 0.0f));
                    }
                    continue;
                }
                // stage the 32 x 32 fp32 block (lane = row) in shared memory, 16-byte groups XOR-swizzled by row, then
                // re-read it with 8 lanes per row so that every global access is a contiguous 128 B (fp32) / 64 B (bf16) row segment
                float4* srow = stage + lane * 8;
#pragma unroll
                for (int gg = 0; gg < 8; gg++)
                    srow[gg ^ (lane & 7)] = make_float4(__uint_as_float(r[4 * gg]), __uint_as_float(r[4 * gg + 1]),
                                                        __uint_as_float(r[4 * gg + 2]), __uint_as_float(r[4 * gg + 3]));
                __syncwarp();
#pragma unroll
                for (int it = 0; it < 8; it++) {
                    const int row = it * 4 + rsub;
                    const int rib = mt * kBM + q * 32 + row;
                    float4 v = stage[row * 8 + (g ^ (row & 7))];
                    if (rib < p.rows_per_batch && col_ok) {
                        v.x += bias4.x; v.y += bias4.y; v.z += bias4.z; v.w += bias4.w;
                        if (EPI == EPI_BIAS_GELU_BF16 || EPI == EPI_BIAS_GELU_POS_F32) {
                            v.x = gelu_tanh(v.x); v.y = gelu_tanh(v.y); v.z = gelu_tanh(v.z); v.w = gelu_tanh(v.w);
                        }
                        const int64_t off = c_base + (int64_t)row * p.ldc + n;
                        if (EPI == EPI_BIAS_GELU_SPLIT) {
                            v.x = gelu_tanh_precise(v.x); v.y = gelu_tanh_precise(v.y); v.z = gelu_tanh_precise(v.z); v.w = gelu_tanh_precise(v.w);
                            const __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y), h1 = __floats2bfloat162_rn(v.z, v.w);
                            const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
                            uint2 whi, wlo;
                            whi.x = *reinterpret_cast<const uint32_t*>(&h0);
                            whi.y = *reinterpret_cast<const uint32_t*>(&h1);
                            wlo.x = pack_bf16(v.x - f0.x, v.y - f0.y);
                            wlo.y = pack_bf16(v.z - f1.x, v.w - f1.y);
                            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + (int64_t)batch * p.c_batch_stride + (int64_t)(mt * kBM + q * 32 + row) * p.ldc + n;
                            *reinterpret_cast<uint2*>(o) = whi;
                            *reinterpret_cast<uint2*>(o + p.split_stride) = wlo;
                        } else if (EPI == EPI_HEADS_BF16) {
                            const int gi = rib / p.group_rows, t = rib - gi * p.group_rows;
                            const int half_n = p.N >> 1, sel = n >= half_n, nn = n - sel * half_n;
                            const int64_t o = ((((int64_t)gi * (half_n >> 6) + (nn >> 6)) * 2 + sel) * p.group_rows + t) * 64 + (nn & 63);
                            uint2 w;
                            w.x = pack_bf16(v.x, v.y);
                            w.y = pack_bf16(v.z, v.w);
                            *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out) + o) = w;
                        } else if (EPI == EPI_BIAS_BF16 || EPI == EPI_BIAS_GELU_BF16 || EPI == EPI_QKV_BF16 || EPI == EPI_BIAS_RELU_BF16 ||
                                   EPI == EPI_BIAS_ADD_RELU_BF16) {
                            if (EPI == EPI_BIAS_ADD_RELU_BF16) {
                                const uint2 rr = rres[it];
                                const float2 r0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&rr.x));
                                const float2 r1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&rr.y));
                                v.x += r0.x; v.y += r0.y; v.z += r1.x; v.w += r1.y;
                            }
                            if (EPI == EPI_BIAS_RELU_BF16 || EPI == EPI_BIAS_ADD_RELU_BF16) {
                                v.x = fmaxf(v.x, 0.0f); v.y = fmaxf(v.y, 0.0f); v.z = fmaxf(v.z, 0.0f); v.w = fmaxf(v.w, 0.0f);
                            }
                            if (CONV != 0 && !((interior >> it) & 1u)) v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                            uint2 w;
                            w.x = pack_bf16(v.x, v.y);
                            w.y = pack_bf16(v.z, v.w);
                            *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out) + off) = w;
                        } else {
                            if (EPI == EPI_BIAS_RESID_F32 || EPI == EPI_BIAS_GELU_POS_F32) {
                                v.x += extra[it].x; v.y += extra[it].y; v.z += extra[it].z; v.w += extra[it].w;
                            }
                            *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + off) = v;
                        }
                    }
                }
                __syncwarp();  // the staging block is reused by the next chunk
            }
            if (!waited) {  // this warp's column half lies beyond N: still take part in the accumulator hand-shake
                mbar_wait(&bar_tfull[as], aph);
                tc_fence_after();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_tempty[as]);
            as ^= 1;
            if (as == 0) aph ^= 1;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = nullptr;
static std::once_flag g_encode_once;

static PFN_encodeTiled get_encode() {
    std::call_once(g_encode_once, [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            g_encode = reinterpret_cast<PFN_encodeTiled>(fn);
    });
    return g_encode;
}

// bf16 tensor map with 128-byte swizzle; dims/strides innermost first; strides in BYTES for dims 1..rank-1.
int make_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return WDR_ERR_CUDA; }
    cuuint64_t gdim[5], gstr[5];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; i++) { gdim[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
    for (int i = 0; i + 1 < rank; i++) gstr[i] = strides_bytes[i];
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu,%llu,%llu] stride0 %llu base %p", (int)r, rank,
                  (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0), (unsigned long long)(rank > 2 ? dims[2] : 0),
                  (unsigned long long)(rank > 1 ? strides_bytes[0] : 0), base);
        return WDR_ERR_CUDA;
    }
    return WDR_OK;
}

static thread_local int g_last_bn = 0;  // tile width the last gemm_bf16 call on this thread picked (tests assert that the 256-wide path ran)
int gemm_last_bn() { return g_last_bn; }

// SM count of the CURRENT device (a process may hold contexts on several devices)
int num_sms() {
    static std::atomic<int> sms[64];
    int dev = 0;
    cudaGetDevice(&dev);
    int v = sms[dev & 63].load(std::memory_order_relaxed);
    if (!v) {
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        if (v <= 0) v = 148;
        sms[dev & 63].store(v, std::memory_order_relaxed);
    }
    return v;
}

template <int BN, int STAGES, int EPI, bool DUAL = false, int CONV = 0>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t st, bool pdl = false) {
    constexpr size_t smem = (size_t)STAGES * ((DUAL ? 2 : 1) * kBM * kBK * 2 + BN * kBK * 2) + 1024;
    static DeviceOnce attr_once;
    WDR_CUDA_TRY(per_device_once(attr_once, [] { return cudaFuncSetAttribute(gemm_bf16_kernel<BN, STAGES, EPI, DUAL, CONV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); }));
    int grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
    WDR_CUDA_TRY(launch_kernel(gemm_bf16_kernel<BN, STAGES, EPI, DUAL, CONV>, dim3(grid), dim3(kGemmThreads), smem, st, pdl, ta, tb, p));
    WDR_LAUNCH_CHECK();
    return WDR_OK;
}

int gemm_bf16(const GemmDesc& d, cudaStream_t st) {
    WDR_REQUIRE(d.A && d.W && d.out && d.rows_per_batch > 0 && d.n_batch > 0 && d.N > 0 && d.K > 0, "bad arguments");
    WDR_REQUIRE(d.N % 8 == 0 && d.K % 8 == 0, "N and K must be multiples of 8");
    WDR_REQUIRE(d.a_row_stride % 8 == 0 && d.a_batch_stride % 8 == 0 && d.ldw % 8 == 0, "operand strides must be multiples of 8 elements");
    WDR_REQUIRE((reinterpret_cast<uintptr_t>(d.A) & 15) == 0 && (reinterpret_cast<uintptr_t>(d.W) & 15) == 0, "operands must be 16-byte aligned");
    WDR_REQUIRE(d.ldc % 8 == 0, "ldc must be a multiple of 8");
    if (d.epilogue == EPI_QKV_BF16) WDR_REQUIRE(d.out_t && d.n_split % 32 == 0, "QKV epilogue needs out_t and n_split % 32 == 0");
    int BN = d.bn == 64 ? 64 : 128;
    // Wide tiles for the big encoder-side GEMMs: a 128 x 128 tile needs 32 KB of operands per 256 tensor-clocks (128 B/clk per SM,
    // at the L2 -> SM limit, measured 64-71 % of the sustained bf16 peak); 128 x 256 needs 48 KB per 512 clocks (96 B/clk).
    static const bool wide_ok = getenv("WDR_GEMM_NO_BN256") == nullptr;
    if (wide_ok && BN == 128 && !d.dual_a && !d.conv2d && d.split_k == 1 && d.N % 256 == 0 &&
        (int64_t)((d.rows_per_batch + kBM - 1) / kBM) * d.n_batch * (d.N / 256) >= 2 * num_sms() &&
        (d.epilogue == EPI_BIAS_BF16 || d.epilogue == EPI_BIAS_GELU_BF16 || d.epilogue == EPI_BIAS_RESID_F32 || d.epilogue == EPI_QKV_BF16 ||
         d.epilogue == EPI_BIAS_GELU_POS_F32 || d.epilogue == EPI_HEADS_BF16))
        BN = 256;
    g_last_bn = BN;
    WDR_REQUIRE(d.split_k >= 1, "split_k must be >= 1");
    if (d.epilogue == EPI_BIAS_GELU_SPLIT) WDR_REQUIRE(d.dual_a && d.split_k == 1 && d.bn == 64 && d.bias && d.split_stride > 0, "EPI_BIAS_GELU_SPLIT is the decoder fc1 GEMM (dual-A, BN=64, no split-K)");
    else if (d.split_k > 1 || (d.bn == 64 && !d.conv2d) || d.dual_a) WDR_REQUIRE(d.epilogue == EPI_F32 && (d.split_k == 1 || !d.bias), "split-K / BN=64 / dual-A are plain fp32-partial GEMMs (EPI_F32, no bias)");
    if (d.conv2d) {
        WDR_REQUIRE(d.conv2d >= 1 && d.conv2d <= 3 && d.tile_pitch && d.tile_row0 && d.conv_F > 0 && d.n_batch == 1 && !d.dual_a && d.split_k == 1,
                    "conv2d GEMMs are single-batch plain-tile GEMMs with per-tile map tables");
        WDR_REQUIRE(d.epilogue == EPI_BIAS_BF16 || d.epilogue == EPI_BIAS_RELU_BF16 || d.epilogue == EPI_BIAS_ADD_RELU_BF16, "conv2d epilogues: bias / bias+ReLU / bias+residual+ReLU");
        WDR_REQUIRE(d.rows_per_batch % kBM == 0, "padded maps are blocks of whole 128-row tiles");
        if (d.conv2d == 1) WDR_REQUIRE(d.a_cols > 0 && d.a_cols % kBK == 0 && d.K == 9 * d.a_cols && d.kb_per_tap == d.a_cols / kBK && d.a_row_stride == d.a_cols, "conv2d = 1: K = 9 C, C a multiple of 64");
        if (d.conv2d == 2) WDR_REQUIRE(d.a_row_stride == 32 && d.K == 6 * kBK, "conv2d = 2: C = 32, K = 3 x 128");
    }
    CUtensorMap ta, tb;
    if (d.conv2d == 1 || d.conv2d == 2) {
        // the activation matrix itself: [rows][C] (C >= 64), or overlapping 64-element rows at a 32-element stride (C = 32: the caller's
        // buffer holds 32 elements of slack behind the last row).  Shifted tiles reach before row 0 / behind the last row: zero fill.
        const uint64_t dims[3] = {(uint64_t)(d.conv2d == 1 ? d.a_cols : 64), (uint64_t)d.rows_per_batch, 1};
        const uint64_t str[2] = {(uint64_t)d.a_row_stride * 2, (uint64_t)d.a_row_stride * 2 * (uint64_t)d.rows_per_batch};
        const uint32_t box[3] = {kBK, kBM, 1};
        int rc = make_tmap_bf16(&ta, d.A, 3, dims, str, box);
        if (rc != WDR_OK) return rc;
    } else {
        // in tap mode the last tap reads rows up to rows_per_batch - 1 + (taps - 1): the caller's buffer holds them
        const int num_kb = (d.K + kBK - 1) / kBK;
        const int taps = d.kb_per_tap > 0 ? (num_kb + d.kb_per_tap - 1) / d.kb_per_tap : 1;
        if (d.dual_a) WDR_REQUIRE(d.n_batch == 1 && (d.bn == 64 || d.bn == 128) && d.a_dual_stride > 0 && d.a_dual_stride % 8 == 0, "dual-A GEMMs are single-batch BN=64/128 GEMMs");
        const uint64_t dims[3] = {(uint64_t)(d.a_cols > 0 ? d.a_cols : d.K), (uint64_t)(d.rows_per_batch + taps - 1), (uint64_t)(d.dual_a ? 2 : d.n_batch)};
        const uint64_t str[2] = {(uint64_t)d.a_row_stride * 2,
                                 (uint64_t)(d.dual_a ? d.a_dual_stride : d.n_batch > 1 ? d.a_batch_stride : d.a_row_stride * d.rows_per_batch) * 2};
        const uint32_t box[3] = {kBK, kBM, 1};
        int rc = make_tmap_bf16(&ta, d.A, 3, dims, str, box);
        if (rc != WDR_OK) return rc;
    }
    if (d.w_kb_major) {
        WDR_REQUIRE(d.K % kBK == 0 && d.ldw == d.K && d.kb_per_tap == 0, "kb-major weights need K %% 64 == 0 and a dense [N][K] source");
        const uint64_t dims[2] = {(uint64_t)kBK, (uint64_t)(d.K / kBK) * (uint64_t)d.N};
        const uint64_t str[1] = {(uint64_t)kBK * 2};
        const uint32_t box[2] = {kBK, (uint32_t)BN};
        int rc = make_tmap_bf16(&tb, d.W, 2, dims, str, box);
        if (rc != WDR_OK) return rc;
    } else {
        const uint64_t dims[2] = {(uint64_t)d.K, (uint64_t)d.N};
        const uint64_t str[1] = {(uint64_t)d.ldw * 2};
        const uint32_t box[2] = {kBK, (uint32_t)BN};
        int rc = make_tmap_bf16(&tb, d.W, 2, dims, str, box);
        if (rc != WDR_OK) return rc;
    }
    GemmParams p;
    p.rows_per_batch = d.rows_per_batch; p.n_batch = d.n_batch; p.N = d.N; p.K = d.K;
    p.kb_per_tap = d.kb_per_tap > 0 ? d.kb_per_tap : (d.K + kBK - 1) / kBK;
    p.tiles_per_batch = (d.rows_per_batch + kBM - 1) / kBM;
    p.n_tiles = (d.N + BN - 1) / BN;
    {
        const int num_kb = (d.K + kBK - 1) / kBK;
        int splits = d.split_k < num_kb ? d.split_k : num_kb;
        p.kb_per_split = (num_kb + splits - 1) / splits;
        p.splits = (num_kb + p.kb_per_split - 1) / p.kb_per_split;  // every split owns at least one k-block
        p.split_stride = d.split_stride;
        if (p.splits != d.split_k && d.split_k > 1) { set_error("gemm: split_k %d does not divide %d k-blocks evenly enough (got %d)", d.split_k, num_kb, p.splits); return WDR_ERR_INVALID; }
    }
    p.total_tiles = p.tiles_per_batch * d.n_batch * p.n_tiles * p.splits;
    p.out = d.out; p.ldc = d.ldc; p.bias = d.bias;
    p.c_batch_stride = d.c_batch_stride > 0 ? d.c_batch_stride : (int64_t)d.rows_per_batch * d.ldc; p.resid = d.resid; p.pos = d.pos;
    p.resid_bf16 = d.resid_bf16;
    p.out_t = d.out_t; p.ldt = d.ldt; p.n_split = d.n_split;
    p.group_rows = d.group_rows;
    p.w_kb_major = d.w_kb_major ? 1 : 0;
    p.tile_pitch = d.tile_pitch; p.tile_row0 = d.tile_row0; p.conv_F = d.conv_F;
    static const int dbg_no_a = getenv("WDR_DEBUG_GEMM_NO_A") ? 1 : 0;
    p.dbg_no_a = dbg_no_a;
    if (d.epilogue == EPI_HEADS_BF16) WDR_REQUIRE(d.group_rows > 0 && d.n_batch == 1 && d.N % 128 == 0, "EPI_HEADS_BF16 needs group_rows, one batch and N = 2 * heads * 64");
    p.t_batch_stride = d.t_batch_stride > 0 ? d.t_batch_stride : d.rows_per_batch;
    if (BN == 256) {
        switch (d.epilogue) {
            case EPI_BIAS_BF16: return launch_gemm<256, 4, EPI_BIAS_BF16>(ta, tb, p, st);
            case EPI_BIAS_GELU_BF16: return launch_gemm<256, 4, EPI_BIAS_GELU_BF16>(ta, tb, p, st);
            case EPI_BIAS_RESID_F32: WDR_REQUIRE(d.resid, "resid missing"); return launch_gemm<256, 4, EPI_BIAS_RESID_F32>(ta, tb, p, st);
            case EPI_BIAS_GELU_POS_F32: WDR_REQUIRE(d.pos, "pos missing"); return launch_gemm<256, 4, EPI_BIAS_GELU_POS_F32>(ta, tb, p, st);
            case EPI_QKV_BF16: return launch_gemm<256, 4, EPI_QKV_BF16>(ta, tb, p, st);
            case EPI_HEADS_BF16: return launch_gemm<256, 4, EPI_HEADS_BF16>(ta, tb, p, st);
            default: break;
        }
    }
    if (d.conv2d && BN == 64) {  // N <= 64: half the B rows and half the MMA width of a 128-wide tile whose upper columns would be padding
#define WDR_CONV64_CASE(E) \
        case E: return d.conv2d == 1 ? launch_gemm<64, 7, E, false, 1>(ta, tb, p, st) : d.conv2d == 2 ? launch_gemm<64, 7, E, false, 2>(ta, tb, p, st) \
                                                                                       : launch_gemm<64, 7, E, false, 3>(ta, tb, p, st);
        switch (d.epilogue) {
            WDR_CONV64_CASE(EPI_BIAS_BF16)
            WDR_CONV64_CASE(EPI_BIAS_RELU_BF16)
            case EPI_BIAS_ADD_RELU_BF16: WDR_REQUIRE(d.resid_bf16, "resid_bf16 missing");
                return d.conv2d == 1 ? launch_gemm<64, 7, EPI_BIAS_ADD_RELU_BF16, false, 1>(ta, tb, p, st) : d.conv2d == 2 ? launch_gemm<64, 7, EPI_BIAS_ADD_RELU_BF16, false, 2>(ta, tb, p, st)
                                                                                                           : launch_gemm<64, 7, EPI_BIAS_ADD_RELU_BF16, false, 3>(ta, tb, p, st);
            default: break;
        }
#undef WDR_CONV64_CASE
    }
    if (d.conv2d) {
#define WDR_CONV_CASE(E) \
        case E: return d.conv2d == 1 ? launch_gemm<128, 5, E, false, 1>(ta, tb, p, st) : d.conv2d == 2 ? launch_gemm<128, 5, E, false, 2>(ta, tb, p, st) \
                                                                                       : launch_gemm<128, 5, E, false, 3>(ta, tb, p, st);
        switch (d.epilogue) {
            WDR_CONV_CASE(EPI_BIAS_BF16)
            WDR_CONV_CASE(EPI_BIAS_RELU_BF16)
            case EPI_BIAS_ADD_RELU_BF16: WDR_REQUIRE(d.resid_bf16, "resid_bf16 missing");
                return d.conv2d == 1 ? launch_gemm<128, 5, EPI_BIAS_ADD_RELU_BF16, false, 1>(ta, tb, p, st) : d.conv2d == 2 ? launch_gemm<128, 5, EPI_BIAS_ADD_RELU_BF16, false, 2>(ta, tb, p, st)
                                                                                                            : launch_gemm<128, 5, EPI_BIAS_ADD_RELU_BF16, false, 3>(ta, tb, p, st);
            default: break;
        }
#undef WDR_CONV_CASE
    }
    switch (d.epilogue) {
        case EPI_BIAS_BF16: return launch_gemm<128, 5, EPI_BIAS_BF16>(ta, tb, p, st);
        case EPI_BIAS_GELU_BF16: return launch_gemm<128, 5, EPI_BIAS_GELU_BF16>(ta, tb, p, st);
        case EPI_BIAS_RESID_F32: WDR_REQUIRE(d.resid, "resid missing"); return launch_gemm<128, 5, EPI_BIAS_RESID_F32>(ta, tb, p, st);
        case EPI_BIAS_GELU_POS_F32: WDR_REQUIRE(d.pos, "pos missing"); return launch_gemm<128, 5, EPI_BIAS_GELU_POS_F32>(ta, tb, p, st);
        case EPI_QKV_BF16: return launch_gemm<128, 5, EPI_QKV_BF16>(ta, tb, p, st);
        case EPI_BIAS_GELU_SPLIT: return launch_gemm<64, 4, EPI_BIAS_GELU_SPLIT, true>(ta, tb, p, st, d.pdl);
        case EPI_HEADS_BF16: return launch_gemm<128, 5, EPI_HEADS_BF16>(ta, tb, p, st);
        case EPI_BIAS_RELU_BF16: return launch_gemm<128, 5, EPI_BIAS_RELU_BF16>(ta, tb, p, st);
        case EPI_BIAS_ADD_RELU_BF16: WDR_REQUIRE(d.resid_bf16, "resid_bf16 missing"); return launch_gemm<128, 5, EPI_BIAS_ADD_RELU_BF16>(ta, tb, p, st);
        case EPI_F32:
            if (d.dual_a) return BN == 64 ? launch_gemm<64, 4, EPI_F32, true>(ta, tb, p, st, d.pdl) : launch_gemm<128, 3, EPI_F32, true>(ta, tb, p, st, d.pdl);
            return BN == 64 ? launch_gemm<64, 7, EPI_F32>(ta, tb, p, st) : launch_gemm<128, 5, EPI_F32>(ta, tb, p, st);
    }
    set_error("unknown epilogue %d", d.epilogue);
    return WDR_ERR_INVALID;
}

}  // namespace wdr
