// embedding.cu — WeSpeaker ResNet34 speaker embeddings on the device: Kaldi fbank -> 2-D ResNet34 (tcgen05 GEMMs) -> TSTP -> Linear.
//
// Replaces pyannote_rs::EmbeddingExtractor::{new, compute} (reference src/transcribe.rs:343, 466-467):
// int16 samples cast to f32 without scaling -> knf fbank (fbank.cu) -> per-column mean subtraction -> ONNX model "feats" [1,T,80]
// -> "embs" [1,D].  The crate ships the CAM++ export; the north-star names WeSpeaker ResNet34: D = 256.
//
// Segments are independent, so a call takes a batch of them.  Activations are bf16, channels innermost.  Level r holds, per segment,
// a ZERO-PADDED map of F = 80 >> r rows and T_r columns stored as packed rows with pitch P = T_r + 1: position (f, t) is row
// (f + 1) * P + t of the segment's block; row f = -1, row f = F and column t = T_r are zero, and a block is a multiple of 128 rows
// (one GEMM M tile never straddles two maps).  In that layout a 3 x 3 stride-1 convolution is NINE SHIFTED GEMMs over the activation
// matrix itself — tap (dy, dx) reads the tile's rows shifted by dy * P + dx — so 29 of the 36 convolutions run as implicit GEMMs on the
// tcgen05 kernel (gemm.cu, conv2d) with no im2col matrix at all; its epilogue (folded BatchNorm shift, residual add, ReLU) writes zero
// at every non-interior position, so every output is again a padded map.  Only the stem (1 input channel), the three stride-2 3 x 3
// convolutions and their 1 x 1 stride-2 shortcuts gather a (small) operand first.  Round 1 materialised the im2col matrix of every
// convolution (9 x the activation bytes, written and read back: 25 + 15 of the 51 ms of a 10 min recording's embedding stage).
// TSTP (mean / unbiased std over time per (channel, frequency)) and the 5120 -> 256 linear layer are fp32.
#include <math.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <vector>
#include "common.cuh"
#include "gemm.cuh"
#include "onnx_file.cuh"
#include "nn_common.cuh"

namespace wdr {

int fbank_run(const int16_t* pcm, const int64_t* seg_offset_dev, const int64_t* feat_offset_dev, const std::vector<int64_t>& seg_offset_host,
              int n_bins, int subtract_mean, float* out, cudaStream_t st, const int64_t* seg_end_dev, const std::vector<int64_t>* seg_end_host);

constexpr int kEmbDimDefault = 256, kEmbBins = 80, kEmbPooled = 5120;  // the width is a property of the loaded model (seg_1 rows): 256 for WeSpeaker ResNet34
constexpr int kEmbMaxFramesPerGroup = 32768;  // fbank frames per forward batch (bounds the im2col workspace: 80 * frames * 288 bf16)

struct ConvW {
    int c_in, c_out, k, stride, K;  // K = GEMM inner size (k*k*c_in, conv1: 16)
    __nv_bfloat16* w;               // [c_out][K], tap-major: column (ky*k + kx)*c_in + c
    float* b;                       // [c_out]
    __nv_bfloat16* w_pair;          // c_in = 32, 3 x 3, stride 1 only: [c_out][3][4][32] — per dy the taps dx = -1, 0, +1 and a zero phantom tap
};
struct BlockW { ConvW c1, c2, sc; bool has_sc; };

static inline uint16_t f32_to_bf16_bits(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    const uint32_t r = ((u >> 16) & 1u) + 0x7FFFu;
    return (uint16_t)((u + r) >> 16);
}

// (segment, f + 1, t) of padded row R: largest seg with poff[seg] <= R, then the position inside that segment's block
__device__ __forceinline__ void emb_locate(const int64_t* __restrict__ poff, const int32_t* __restrict__ T, int n_seg, int64_t R, int& seg, int& f1, int& t, int& Ts) {
    int lo = 0, hi = n_seg - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(poff + mid) <= R) lo = mid; else hi = mid - 1;
    }
    seg = lo;
    Ts = __ldg(T + seg);
    const int local = (int)(R - __ldg(poff + seg)), P = Ts + 1;
    f1 = local / P;
    t = local - f1 * P;
}

// fbank feats [t][80] fp32 -> the stem's operand [padded rows of level 0][16] bf16 (taps (ky, kx) over (f, t); columns 9..15 and the
// non-interior rows zero)
__global__ void emb_im2col_feats_kernel(const float* __restrict__ feats, const int64_t* __restrict__ feat_off, const int32_t* __restrict__ T,
                                        const int64_t* __restrict__ poff, int n_seg, int64_t total_rows, __nv_bfloat16* __restrict__ A) {
    for (int64_t R = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; R < total_rows; R += (int64_t)gridDim.x * blockDim.x) {
        int seg, f1, t, Ts;
        emb_locate(poff, T, n_seg, R, seg, f1, t, Ts);
        const int f = f1 - 1;
        __align__(16) __nv_bfloat16 v[16];
#pragma unroll
        for (int tap = 0; tap < 16; tap++) v[tap] = __float2bfloat16_rn(0.0f);
        if (f >= 0 && f < kEmbBins && t < Ts) {
            const float* x = feats + __ldg(feat_off + seg) * kEmbBins;
#pragma unroll
            for (int tap = 0; tap < 9; tap++) {
                const int fi = f + tap / 3 - 1, ti = t + tap % 3 - 1;
                if (fi >= 0 && fi < kEmbBins && ti >= 0 && ti < Ts) v[tap] = __float2bfloat16_rn(x[(int64_t)ti * kEmbBins + fi]);
            }
        }
        uint4* dst = reinterpret_cast<uint4*>(A + R * 16);
        dst[0] = *reinterpret_cast<const uint4*>(v);
        dst[1] = *reinterpret_cast<const uint4*>(v + 8);
    }
}

// Stride-2 convolutions (3 x 3 and the 1 x 1 shortcut): padded map of level r -> operand [padded rows of level r + 1][KS*KS*C] bf16.
// One WARP per output row (a binary search over the block offsets finds its segment), lanes over the row's 16-byte vectors (8
// channels of one tap); rows that are not interior positions of the output map are zero.
template <int KS>
__global__ void __launch_bounds__(256)
emb_im2col_s2_kernel(const __nv_bfloat16* __restrict__ in, int C, int F_in, const int32_t* __restrict__ T_in, const int32_t* __restrict__ T_out,
                     const int64_t* __restrict__ in_poff, const int64_t* __restrict__ out_poff, int n_seg, int64_t total_rows, __nv_bfloat16* __restrict__ A) {
    const int lane = threadIdx.x & 31;
    const int cv_n = C >> 3, cv_shift = 31 - __clz(cv_n), vec_per_row = KS * KS * cv_n;  // C is a power of two (32 .. 128)
    const int F_out = F_in >> 1;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t R = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); R < total_rows; R += warps) {
        int seg, f1, t, To;
        emb_locate(out_poff, T_out, n_seg, R, seg, f1, t, To);
        const int f = f1 - 1;
        const bool interior = f >= 0 && f < F_out && t < To;
        const int Ti = __ldg(T_in + seg), Pi = Ti + 1;
        const uint4* src = reinterpret_cast<const uint4*>(in) + __ldg(in_poff + seg) * cv_n;
        uint4* dst = reinterpret_cast<uint4*>(A) + R * vec_per_row;
        for (int v = lane; v < vec_per_row; v += 32) {
            const int tap = v >> cv_shift, cv = v & (cv_n - 1);
            const int ky = tap / KS, kx = tap - ky * KS;
            const int fi = f * 2 + ky - KS / 2, ti = t * 2 + kx - KS / 2;
            uint4 val = make_uint4(0, 0, 0, 0);
            if (interior && fi >= 0 && fi < F_in && ti >= 0 && ti < Ti) val = __ldg(src + ((int64_t)(fi + 1) * Pi + ti) * cv_n + cv);
            dst[v] = val;
        }
    }
}

// TSTP: x = padded maps of level 3 (F = 10, 256 channels) -> stats[seg][c*10 + f] = mean_t, stats[seg][2560 + c*10 + f] = sqrt(unbiased var_t + 1e-7)
__global__ void emb_tstp_kernel(const __nv_bfloat16* __restrict__ x, const int32_t* __restrict__ T3, const int64_t* __restrict__ poff3,
                                float* __restrict__ stats) {
    const int seg = blockIdx.y, f = blockIdx.x, c = threadIdx.x;  // 256 threads = channels
    const int Ts = T3[seg];
    const __nv_bfloat16* p = x + (poff3[seg] + (int64_t)(f + 1) * (Ts + 1)) * 256 + c;
    double s = 0.0, q = 0.0;
    for (int t = 0; t < Ts; t++) {
        const double v = (double)__bfloat162float(p[(int64_t)t * 256]);
        s += v; q += v * v;
    }
    const double mean = s / Ts;
    double var = (q - s * mean) / (Ts > 1 ? Ts - 1 : 1);
    if (var < 0.0) var = 0.0;
    stats[(int64_t)seg * kEmbPooled + c * 10 + f] = (float)mean;
    stats[(int64_t)seg * kEmbPooled + 2560 + c * 10 + f] = (float)sqrt(var + 1e-7);
}

}  // namespace wdr

namespace wdr { struct EmbProf; }
using namespace wdr;

struct wdr_emb {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::vector<void*> allocs;
    ConvW conv1{};
    std::vector<BlockW> blocks;
    float *lin_w = nullptr, *lin_b = nullptr;
    // workspaces (grown on demand)
    __nv_bfloat16 *act[4] = {nullptr, nullptr, nullptr, nullptr}, *col = nullptr;
    size_t act_cap = 0, col_cap = 0;
    int32_t* tiles = nullptr;  // per level and 128-row tile: pitch of the map that owns it, row where that map's block starts
    size_t tiles_cap = 0;
    wdr::DevArena io;       // host-pointer API: staged PCM + result embeddings
    wdr::DevArena scratch;  // per-group features, level tables, pooled statistics, embeddings (grow-only)
    double conv_flops = 0.0;  // of the last call (algorithmic, 2*M*N*K)
    wdr::EmbProf* prof = nullptr;          // wdr_emb_profile
    double gemm_ms = 0.0, gather_ms = 0.0; // of the last profiled call
    int emb_dim = wdr::kEmbDimDefault;  // rows of the embedding layer (runtime: read from the ONNX file when one is loaded)
};

namespace wdr {

// file != nullptr: the convolution comes from an ONNX export with its BatchNorm already folded (onnx_extract_resnet34): weight
// [co][ci][k][k] and bias are used as they are; otherwise seeded tensors are drawn and folded here.
// c_in = 32, 3 x 3, stride 1: the weights once more as [c_out][dy][4 taps][32] (dx = -1, 0, +1 and a zero phantom tap), the K layout of
// the implicit GEMM over overlapping 64-element activation rows (gemm.cuh, conv2d = 2)
static int emb_upload_pair_layout(wdr_emb* m, const std::vector<uint16_t>& wb, int ci, int co, int k, int stride, int K, __nv_bfloat16** out) {
    *out = nullptr;
    if (!(ci == 32 && k == 3 && stride == 1)) return WDR_OK;
    std::vector<uint16_t> wp((size_t)co * 384, 0);
    for (int o = 0; o < co; o++)
        for (int ky = 0; ky < 3; ky++)
            for (int kx = 0; kx < 3; kx++)
                for (int c = 0; c < 32; c++) wp[(size_t)o * 384 + (size_t)(ky * 4 + kx) * 32 + c] = wb[(size_t)o * K + (size_t)(ky * 3 + kx) * 32 + c];
    WDR_CUDA_TRY(cudaMalloc(out, sizeof(uint16_t) * wp.size()));
    m->allocs.push_back(*out);
    WDR_CUDA_TRY(cudaMemcpy(*out, wp.data(), sizeof(uint16_t) * wp.size(), cudaMemcpyHostToDevice));
    return WDR_OK;
}

static int emb_upload_conv(wdr_emb* m, uint64_t seed, const std::string& name, int ci, int co, int k, int stride, ConvW* out, const NamedTensors* file = nullptr) {
    const int fan = ci * k * k;
    const std::string base = "resnet34." + name;
    if (file) {
        auto iw = file->find(name + ".weight"), ib = file->find(name + ".bias");
        if (iw == file->end() || ib == file->end() || iw->second.size() != (size_t)co * fan || ib->second.size() != (size_t)co) { set_error("wdr_emb_init: %s missing from the model file", name.c_str()); return WDR_ERR_INVALID; }
        const int K = ci == 1 ? 16 : fan;
        std::vector<uint16_t> wb((size_t)co * K, 0);
        for (int o = 0; o < co; o++)
            for (int c = 0; c < ci; c++)
                for (int t = 0; t < k * k; t++) wb[(size_t)o * K + (size_t)t * ci + c] = f32_to_bf16_bits(iw->second[((size_t)o * ci + c) * k * k + t]);
        __nv_bfloat16* dw = nullptr;
        float* db = nullptr;
        WDR_CUDA_TRY(cudaMalloc(&dw, sizeof(uint16_t) * wb.size()));
        m->allocs.push_back(dw);
        WDR_CUDA_TRY(cudaMalloc(&db, sizeof(float) * co));
        m->allocs.push_back(db);
        WDR_CUDA_TRY(cudaMemcpy(dw, wb.data(), sizeof(uint16_t) * wb.size(), cudaMemcpyHostToDevice));
        WDR_CUDA_TRY(cudaMemcpy(db, ib->second.data(), sizeof(float) * co, cudaMemcpyHostToDevice));
        __nv_bfloat16* dp = nullptr;
        const int rcp = emb_upload_pair_layout(m, wb, ci, co, k, stride, K, &dp);
        if (rcp != WDR_OK) return rcp;
        *out = ConvW{ci, co, k, stride, K, dw, db, dp};
        return WDR_OK;
    }
    std::vector<float> w = nn_synth(seed, base + ".weight", (size_t)co * fan, 0.0f, (float)sqrt(6.0 / fan));
    const bool tail = name.size() >= 5 && (name.compare(name.size() - 5, 5, "conv2") == 0 || name.compare(name.size() - 8 > name.size() ? 0 : name.size() - 8, 8, "shortcut") == 0);
    std::vector<float> g = nn_synth(seed, base + ".bn.weight", co, tail ? 0.7f : 1.0f, 0.1f);
    std::vector<float> beta = nn_synth(seed, base + ".bn.bias", co, 0.0f, 0.1f);
    std::vector<float> mean = nn_synth(seed, base + ".bn.running_mean", co, 0.0f, 0.1f);
    std::vector<float> var = nn_synth(seed, base + ".bn.running_var", co, 1.0f, 0.2f);
    const int K = ci == 1 ? 16 : fan;
    std::vector<uint16_t> wb((size_t)co * K, 0);
    std::vector<float> bias(co);
    for (int o = 0; o < co; o++) {
        volatile float denom = sqrtf(var[o] + 1e-5f);
        volatile float s = g[o] / denom;
        volatile float ms = mean[o] * s;
        bias[o] = beta[o] - ms;
        // synthetic tensor layout is PyTorch's [co][ci][ky][kx]; the GEMM wants [co][(ky*k + kx)*ci + c]
        for (int c = 0; c < ci; c++)
            for (int t = 0; t < k * k; t++) {
                volatile float prod = w[((size_t)o * ci + c) * k * k + t] * s;
                wb[(size_t)o * K + (size_t)t * ci + c] = f32_to_bf16_bits(prod);
            }
    }
    __nv_bfloat16* dw = nullptr;
    float* db = nullptr;
    WDR_CUDA_TRY(cudaMalloc(&dw, sizeof(uint16_t) * wb.size()));
    m->allocs.push_back(dw);
    WDR_CUDA_TRY(cudaMalloc(&db, sizeof(float) * co));
    m->allocs.push_back(db);
    WDR_CUDA_TRY(cudaMemcpy(dw, wb.data(), sizeof(uint16_t) * wb.size(), cudaMemcpyHostToDevice));
    WDR_CUDA_TRY(cudaMemcpy(db, bias.data(), sizeof(float) * co, cudaMemcpyHostToDevice));
    __nv_bfloat16* dp = nullptr;
    const int rcp = emb_upload_pair_layout(m, wb, ci, co, k, stride, K, &dp);
    if (rcp != WDR_OK) return rcp;
    *out = ConvW{ci, co, k, stride, K, dw, db, dp};
    return WDR_OK;
}

// Measurement aid (wdr_emb_profile): a CUDA-event pair on the launching stream around every GEMM / every im2col gather of a call,
// so that bench.py can quote the tcgen05 GEMMs of this stage against the tensor roofline without the gathers, fbank and copies.
struct EmbProf {
    bool on = false;
    std::vector<cudaEvent_t> pool;
    size_t used = 0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> gemm, gather;
    cudaEvent_t get() {
        if (used == pool.size()) { cudaEvent_t e; cudaEventCreate(&e); pool.push_back(e); }
        return pool[used++];
    }
};
static EmbProf* g_emb_prof = nullptr;  // set for the duration of one emb_forward (single caller per model, &mut self upstream)
struct EmbProfScope {
    cudaEvent_t a = nullptr, b = nullptr;
    cudaStream_t st;
    bool gemm;
    EmbProfScope(bool is_gemm, cudaStream_t s) : st(s), gemm(is_gemm) {
        if (g_emb_prof && g_emb_prof->on) { a = g_emb_prof->get(); b = g_emb_prof->get(); cudaEventRecord(a, st); }
    }
    ~EmbProfScope() {
        if (a) { cudaEventRecord(b, st); (gemm ? g_emb_prof->gemm : g_emb_prof->gather).push_back({a, b}); }
    }
};

struct EmbLevel {
    int F;
    std::vector<int32_t> T;
    std::vector<int64_t> off;  // n + 1: block offsets of the zero-padded maps (multiples of 128 rows)
    int64_t interior = 0;      // sum of F * T: the rows a convolution really computes (FLOP accounting)
    int32_t* d_T = nullptr;
    int64_t* d_off = nullptr;
    int32_t *d_pitch = nullptr, *d_row0 = nullptr;  // per 128-row tile
    int64_t rows() const { return off.back(); }
};

// One convolution as a GEMM.  A = nullptr: implicit (3 x 3, stride 1: the operand is the padded input map `in` itself); otherwise A is
// a gathered operand [rows][c.K].  The output is a padded map of level `lv`.
static int emb_gemm(const __nv_bfloat16* A, const __nv_bfloat16* in, const EmbLevel& lv, const ConvW& c, int epi, const __nv_bfloat16* resid, __nv_bfloat16* out,
                    cudaStream_t st) {
    GemmDesc g;
    g.rows_per_batch = (int)lv.rows(); g.n_batch = 1;
    g.N = c.c_out;
    g.epilogue = epi; g.out = out; g.ldc = c.c_out; g.bias = c.b; g.resid_bf16 = resid;
    g.tile_pitch = lv.d_pitch; g.tile_row0 = lv.d_row0; g.conv_F = lv.F;
    g.bn = c.c_out <= 64 ? 64 : 128;
    if (A) {
        g.A = A; g.a_row_stride = c.K; g.W = c.w; g.ldw = c.K; g.K = c.K; g.conv2d = 3;
    } else if (c.c_in == 32) {
        g.A = in; g.a_row_stride = 32; g.W = c.w_pair; g.ldw = 384; g.K = 384; g.conv2d = 2;
    } else {
        g.A = in; g.a_row_stride = c.c_in; g.a_cols = c.c_in; g.kb_per_tap = c.c_in / 64; g.W = c.w; g.ldw = c.K; g.K = c.K; g.conv2d = 1;
    }
    EmbProfScope ps(true, st);
    return gemm_bf16(g, st);
}

template <typename T>
static int emb_grow(T** p, size_t* cap, size_t need) {
    if (need <= *cap) return WDR_OK;
    if (*p) cudaFree(*p);
    *p = nullptr; *cap = 0;
    WDR_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(p), need * sizeof(T)));
    *cap = need;
    return WDR_OK;
}

// One forward batch: feats (device, [frames][80], segment s at frame feat_off[s]) -> out_dev[n][256]
static int emb_forward(wdr_emb* m, const float* feats, const std::vector<int64_t>& feat_off, float* out_dev, int32_t* tab_T, int64_t* tab_off,
                       float* stats_buf, cudaStream_t st) {
    const int n = (int)feat_off.size() - 1;
    static const int chan[4] = {32, 64, 128, 256};
    EmbLevel lv[4];
    size_t n_tiles_all = 0;
    for (int r = 0; r < 4; r++) {
        lv[r].F = kEmbBins >> r;
        lv[r].T.resize(n);
        lv[r].off.assign(n + 1, 0);
        for (int s = 0; s < n; s++) {
            const int T0 = (int)(feat_off[s + 1] - feat_off[s]);
            int T = T0;
            for (int k = 0; k < r; k++) T = (T + 1) / 2;
            lv[r].T[s] = T;
            const int64_t block = ((int64_t)(lv[r].F + 2) * (T + 1) + 127) / 128 * 128;
            lv[r].off[s + 1] = lv[r].off[s] + block;
            lv[r].interior += (int64_t)lv[r].F * T;
        }
        n_tiles_all += (size_t)(lv[r].rows() / 128);
    }
    WDR_REQUIRE(lv[0].rows() < (int64_t)1 << 31, "embedding batch too large");
    int rc;
    if ((rc = emb_grow(&m->tiles, &m->tiles_cap, 2 * n_tiles_all)) != WDR_OK) return rc;
    // level tables + feat offsets on the device
    std::vector<int32_t> tiles_host(2 * n_tiles_all);
    {
        size_t at = 0;
        for (int r = 0; r < 4; r++) {
            const size_t nt = (size_t)(lv[r].rows() / 128);
            lv[r].d_pitch = m->tiles + at;
            lv[r].d_row0 = m->tiles + at + nt;
            for (int s = 0; s < n; s++)
                for (int64_t tl = lv[r].off[s] / 128; tl < lv[r].off[s + 1] / 128; tl++) {
                    tiles_host[at + (size_t)tl] = lv[r].T[s] + 1;
                    tiles_host[at + nt + (size_t)tl] = (int32_t)lv[r].off[s];
                }
            at += 2 * nt;
        }
    }
    WDR_CUDA_TRY(cudaMemcpyAsync(m->tiles, tiles_host.data(), sizeof(int32_t) * tiles_host.size(), cudaMemcpyHostToDevice, st));
    for (int r = 0; r < 4; r++) {
        lv[r].d_T = tab_T + (size_t)r * n;
        lv[r].d_off = tab_off + (size_t)r * (n + 1);
        WDR_CUDA_TRY(cudaMemcpyAsync(lv[r].d_T, lv[r].T.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
        WDR_CUDA_TRY(cudaMemcpyAsync(lv[r].d_off, lv[r].off.data(), sizeof(int64_t) * (n + 1), cudaMemcpyHostToDevice, st));
    }
    int64_t* d_feat_off = tab_off + (size_t)4 * (n + 1);
    WDR_CUDA_TRY(cudaMemcpyAsync(d_feat_off, feat_off.data(), sizeof(int64_t) * (n + 1), cudaMemcpyHostToDevice, st));
    // workspaces: activations of the widest level (+ slack: the C = 32 implicit GEMM views rows 64 elements wide); the gathered operands
    // of the stem (16 columns) and of the stride-2 convolutions (9 C_in columns at the OUTPUT level's rows)
    size_t act_need = 0, col_need = (size_t)lv[0].rows() * 16;
    for (int r = 0; r < 4; r++) act_need = std::max(act_need, (size_t)lv[r].rows() * chan[r]);
    for (int r = 1; r < 4; r++) col_need = std::max(col_need, (size_t)lv[r].rows() * 9 * chan[r - 1]);
    act_need += 1024; col_need += 1024;
    if (act_need > m->act_cap) {
        for (int i = 0; i < 4; i++) { if (m->act[i]) cudaFree(m->act[i]); m->act[i] = nullptr; }
        m->act_cap = 0;
        for (int i = 0; i < 4; i++) {
            WDR_CUDA_TRY(cudaMalloc(&m->act[i], sizeof(__nv_bfloat16) * act_need));
            // every map is written whole before it is read, except the 32 elements behind the last row that the C = 32 implicit GEMM's
            // 64-wide row view touches (against zero weights): they must hold finite values
            WDR_CUDA_TRY(cudaMemsetAsync(m->act[i], 0, sizeof(__nv_bfloat16) * act_need, st));
        }
        m->act_cap = act_need;
    }
    if ((rc = emb_grow(&m->col, &m->col_cap, col_need)) != WDR_OK) return rc;
    g_emb_prof = m->prof;
    double flops = 0.0;
    // stem
    {
        {
            EmbProfScope ps(false, st);
            const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>((lv[0].rows() + 255) / 256, 148 * 16));
            emb_im2col_feats_kernel<<<grid, 256, 0, st>>>(feats, d_feat_off, lv[0].d_T, lv[0].d_off, n, lv[0].rows(), m->col);
        }
        WDR_LAUNCH_CHECK();
        if ((rc = emb_gemm(m->col, nullptr, lv[0], m->conv1, EPI_BIAS_RELU_BF16, nullptr, m->act[0], st)) != WDR_OK) return rc;
        flops += 2.0 * lv[0].interior * 32 * 9;
    }
    __nv_bfloat16 *x = m->act[0], *y1 = m->act[1], *sc = m->act[2], *x2 = m->act[3];
    int level = 0;
    // a convolution of input map `in` (level lin) into a padded map of level lout
    auto conv = [&](const __nv_bfloat16* in, const ConvW& c, int lin, int lout, int epi, const __nv_bfloat16* resid, __nv_bfloat16* out) -> int {
        flops += 2.0 * lv[lout].interior * c.c_out * c.c_in * c.k * c.k;
        if (c.k == 3 && c.stride == 1) return emb_gemm(nullptr, in, lv[lout], c, epi, resid, out, st);   // implicit GEMM
        if (c.k == 1 && c.stride == 1) return emb_gemm(in, nullptr, lv[lout], c, epi, resid, out, st);   // the map itself is the operand
        {
            const int64_t total_rows = lv[lout].rows();
            const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>((total_rows + 7) / 8, 148 * 16));
            EmbProfScope ps(false, st);
            if (c.k == 3)
                emb_im2col_s2_kernel<3><<<grid, 256, 0, st>>>(in, c.c_in, lv[lin].F, lv[lin].d_T, lv[lout].d_T, lv[lin].d_off, lv[lout].d_off, n, total_rows, m->col);
            else
                emb_im2col_s2_kernel<1><<<grid, 256, 0, st>>>(in, c.c_in, lv[lin].F, lv[lin].d_T, lv[lout].d_T, lv[lin].d_off, lv[lout].d_off, n, total_rows, m->col);
            WDR_LAUNCH_CHECK();
        }
        return emb_gemm(m->col, nullptr, lv[lout], c, epi, resid, out, st);
    };
    for (const BlockW& b : m->blocks) {
        const int lout = level + (b.c1.stride == 2 ? 1 : 0);
        WDR_REQUIRE(b.c1.stride == 1 || b.c1.stride == 2, "ResNet34 strides are 1 or 2");
        if ((rc = conv(x, b.c1, level, lout, EPI_BIAS_RELU_BF16, nullptr, y1)) != WDR_OK) return rc;
        const __nv_bfloat16* resid = x;
        if (b.has_sc) {
            if ((rc = conv(x, b.sc, level, lout, EPI_BIAS_BF16, nullptr, sc)) != WDR_OK) return rc;
            resid = sc;
        }
        if ((rc = conv(y1, b.c2, lout, lout, EPI_BIAS_ADD_RELU_BF16, resid, x2)) != WDR_OK) return rc;
        std::swap(x, x2);
        level = lout;
    }
    emb_tstp_kernel<<<dim3(10, n), 256, 0, st>>>(x, lv[3].d_T, lv[3].d_off, stats_buf);
    WDR_LAUNCH_CHECK();
    if ((rc = sgemm_nt(stats_buf, kEmbPooled, m->lin_w, kEmbPooled, m->lin_b, out_dev, m->emb_dim, n, m->emb_dim, kEmbPooled, NN_ACT_NONE, st)) != WDR_OK) return rc;
    WDR_CUDA_TRY(cudaStreamSynchronize(st));  // the host-side level tables (async H2D sources) die with this frame
    g_emb_prof = nullptr;
    if (m->prof && m->prof->on) {
        auto total = [](std::vector<std::pair<cudaEvent_t, cudaEvent_t>>& v) {
            double ms = 0.0;
            for (auto& e : v) { float t = 0.0f; if (cudaEventElapsedTime(&t, e.first, e.second) == cudaSuccess) ms += t; }
            v.clear();
            return ms;
        };
        m->gemm_ms += total(m->prof->gemm);
        m->gather_ms += total(m->prof->gather);
        m->prof->used = 0;
    }
    m->conv_flops += flops;
    return WDR_OK;
}

}  // namespace wdr

extern "C" wdr_emb* wdr_emb_init(const char* path, uint64_t seed, int device) {
    clear_error();
    // EmbeddingExtractor::new(path) (src/transcribe.rs:343): a WeSpeaker ResNet34 ONNX export; NULL / "" = seeded weights
    NamedTensors file_w;
    const NamedTensors* fw = nullptr;
    int dim = kEmbDimDefault;
    if (path && path[0]) {
        OnnxFile of;
        std::string err;
        if (!of.load(path, &err) || !onnx_extract_resnet34(of, &file_w, &dim, &err)) { set_error("wdr_emb_init: %s", err.c_str()); return nullptr; }
        if (dim <= 0 || dim > 4096) { set_error("wdr_emb_init: implausible embedding width %d", dim); return nullptr; }
        fw = &file_w;
    }
    if (ensure_device(device) != WDR_OK) return nullptr;
    wdr_emb* m = new wdr_emb();
    m->device = device;
    m->emb_dim = dim;
    if (cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking) != cudaSuccess) { set_error("wdr_emb_init: stream"); delete m; return nullptr; }
    int rc = emb_upload_conv(m, seed, "conv1", 1, 32, 3, 1, &m->conv1, fw);
    static const int planes[4] = {32, 64, 128, 256}, nblocks[4] = {3, 4, 6, 3}, strides[4] = {1, 2, 2, 2};
    int c_in = 32;
    for (int li = 0; li < 4 && rc == WDR_OK; li++)
        for (int bi = 0; bi < nblocks[li] && rc == WDR_OK; bi++) {
            const int s = bi == 0 ? strides[li] : 1;
            char nm[64];
            BlockW b{};
            snprintf(nm, sizeof(nm), "layer%d.%d.conv1", li + 1, bi);
            rc = emb_upload_conv(m, seed, nm, c_in, planes[li], 3, s, &b.c1, fw);
            snprintf(nm, sizeof(nm), "layer%d.%d.conv2", li + 1, bi);
            if (rc == WDR_OK) rc = emb_upload_conv(m, seed, nm, planes[li], planes[li], 3, 1, &b.c2, fw);
            b.has_sc = s != 1 || c_in != planes[li];
            if (b.has_sc && rc == WDR_OK) {
                snprintf(nm, sizeof(nm), "layer%d.%d.shortcut", li + 1, bi);
                rc = emb_upload_conv(m, seed, nm, c_in, planes[li], 1, s, &b.sc, fw);
            }
            m->blocks.push_back(b);
            c_in = planes[li];
        }
    if (rc == WDR_OK) {
        std::vector<float> lw = fw ? file_w["seg_1.weight"] : nn_synth(seed, "resnet34.seg_1.weight", (size_t)dim * kEmbPooled, 0.0f, (float)(1.0 / sqrt(5120.0)));
        std::vector<float> lb = fw ? file_w["seg_1.bias"] : nn_synth(seed, "resnet34.seg_1.bias", dim, 0.0f, 0.05f);
        if (lw.size() != (size_t)dim * kEmbPooled || lb.size() != (size_t)dim) { set_error("wdr_emb_init: embedding layer shape"); rc = WDR_ERR_INVALID; }
        else if (cudaMalloc(&m->lin_w, sizeof(float) * lw.size()) != cudaSuccess || cudaMalloc(&m->lin_b, sizeof(float) * lb.size()) != cudaSuccess) rc = WDR_ERR_CUDA;
        else {
            cudaMemcpy(m->lin_w, lw.data(), sizeof(float) * lw.size(), cudaMemcpyHostToDevice);
            cudaMemcpy(m->lin_b, lb.data(), sizeof(float) * lb.size(), cudaMemcpyHostToDevice);
        }
    }
    if (rc != WDR_OK) {
        if (!wdr_last_error()[0]) set_error("wdr_emb_init: device allocation failed");
        wdr_emb_free(m);
        return nullptr;
    }
    return m;
}

extern "C" void wdr_emb_free(wdr_emb* m) {
    if (!m) return;
    cudaSetDevice(m->device);
    cudaDeviceSynchronize();
    for (void* p : m->allocs) cudaFree(p);
    cudaFree(m->lin_w);
    cudaFree(m->lin_b);
    for (int i = 0; i < 4; i++) cudaFree(m->act[i]);
    cudaFree(m->col);
    cudaFree(m->tiles);
    m->scratch.release();
    m->io.release();
    if (m->stream) cudaStreamDestroy(m->stream);
    if (m->prof) {
        for (auto e : m->prof->pool) cudaEventDestroy(e);
        delete m->prof;
    }
    delete m;
}

extern "C" int wdr_emb_dim(wdr_emb* m) { return m ? m->emb_dim : 0; }

extern "C" double wdr_emb_last_flops(wdr_emb* m) { return m ? m->conv_flops : 0.0; }
extern "C" int wdr_emb_profile(wdr_emb* m, int enable) {
    clear_error();
    WDR_REQUIRE(m, "bad arguments");
    if (!m->prof) m->prof = new EmbProf();
    m->prof->on = enable != 0;
    return WDR_OK;
}
extern "C" int wdr_emb_last_kernel_ms(wdr_emb* m, double* gemm_ms, double* gather_ms) {
    clear_error();
    WDR_REQUIRE(m && gemm_ms && gather_ms, "bad arguments");
    *gemm_ms = m->gemm_ms;
    *gather_ms = m->gather_ms;
    return WDR_OK;
}

// pcm_dev: device int16, segment s = [seg_off[s], seg_off[s+1]) (host offsets).  out_dev [n][256]; status[s] = 0 or WDR_ERR_TOO_SHORT.
static int emb_compute_dev(wdr_emb* m, const int16_t* pcm_dev, const std::vector<int64_t>& seg_off, float* out_dev, int32_t* status, cudaStream_t st) {
    const int n = (int)seg_off.size() - 1;
    m->conv_flops = 0.0;
    m->gemm_ms = m->gather_ms = 0.0;
    // segments that yield at least one frame, in order, cut into forward groups by total frames
    std::vector<int> live;
    for (int s = 0; s < n; s++) {
        const int64_t len = seg_off[s + 1] - seg_off[s];
        WDR_REQUIRE(len >= 0 && len < ((int64_t)1 << 31), "segment length out of range");
        const int T = wdr_fbank_frames((int)len);
        if (status) status[s] = T > 0 ? WDR_OK : WDR_ERR_TOO_SHORT;
        if (T > 0) live.push_back(s);
    }
    size_t i0 = 0;
    while (i0 < live.size()) {
        size_t i1 = i0;
        int64_t frames = 0;
        while (i1 < live.size()) {
            const int T = wdr_fbank_frames((int)(seg_off[live[i1] + 1] - seg_off[live[i1]]));
            if (i1 > i0 && frames + T > kEmbMaxFramesPerGroup) break;
            frames += T;
            i1++;
        }
        const int g = (int)(i1 - i0);
        // this group's segments are generally not contiguous in pcm (short ones are skipped): per-segment offsets for the fbank
        std::vector<int64_t> so(2 * (size_t)g), fo((size_t)g + 1, 0);
        bool contiguous = true;
        for (int k = 0; k < g; k++) {
            const int s = live[i0 + k];
            if (k > 0 && seg_off[s] != seg_off[live[i0 + k - 1] + 1]) contiguous = false;
            fo[k + 1] = fo[k] + wdr_fbank_frames((int)(seg_off[s + 1] - seg_off[s]));
        }
        struct { float* p; } feats, emb;
        struct { int64_t* p; } d_so, d_fo, d_se;
        int rc;
        {
            DevArena& A = m->scratch;
            const size_t need = DevArena::padded(sizeof(float) * frames * kEmbBins) + DevArena::padded(sizeof(float) * g * m->emb_dim) +
                                3 * DevArena::padded(sizeof(int64_t) * (g + 2)) + DevArena::padded(sizeof(int32_t) * 4 * g) +
                                DevArena::padded(sizeof(int64_t) * 5 * (g + 1)) + DevArena::padded(sizeof(float) * g * kEmbPooled);
            if ((rc = A.reserve(need)) != WDR_OK) return rc;
            feats.p = A.take<float>((size_t)frames * kEmbBins);
            emb.p = A.take<float>((size_t)g * m->emb_dim);
            d_fo.p = A.take<int64_t>((size_t)g + 2);
            d_so.p = A.take<int64_t>((size_t)g + 2);
            d_se.p = A.take<int64_t>((size_t)g + 2);
        }
        WDR_CUDA_TRY(cudaMemcpyAsync(d_fo.p, fo.data(), sizeof(int64_t) * (g + 1), cudaMemcpyHostToDevice, st));
        if (contiguous) {
            std::vector<int64_t> s2((size_t)g + 1);
            for (int k = 0; k < g; k++) s2[k] = seg_off[live[i0 + k]];
            s2[g] = seg_off[live[i0 + g - 1] + 1];
            WDR_CUDA_TRY(cudaMemcpyAsync(d_so.p, s2.data(), sizeof(int64_t) * (g + 1), cudaMemcpyHostToDevice, st));
            WDR_CUDA_TRY(cudaStreamSynchronize(st));
            rc = fbank_run(pcm_dev, d_so.p, d_fo.p, s2, kEmbBins, 1, feats.p, st, nullptr, nullptr);
            if (rc != WDR_OK) return rc;
        } else {
            // segments that are not back to back in pcm (the short ones in between were skipped): explicit [start, end) pairs, ONE launch
            // (round 1 launched the fbank once per contiguous run — 39 launch pairs and stream synchronisations per 10 min recording)
            std::vector<int64_t> s2((size_t)g), e2((size_t)g);
            for (int k = 0; k < g; k++) { s2[k] = seg_off[live[i0 + k]]; e2[k] = seg_off[live[i0 + k] + 1]; }
            WDR_CUDA_TRY(cudaMemcpyAsync(d_so.p, s2.data(), sizeof(int64_t) * g, cudaMemcpyHostToDevice, st));
            WDR_CUDA_TRY(cudaMemcpyAsync(d_se.p, e2.data(), sizeof(int64_t) * g, cudaMemcpyHostToDevice, st));
            WDR_CUDA_TRY(cudaStreamSynchronize(st));
            rc = fbank_run(pcm_dev, d_so.p, d_fo.p, s2, kEmbBins, 1, feats.p, st, d_se.p, &e2);
            if (rc != WDR_OK) return rc;
        }
        {
            int32_t* tab_T = m->scratch.take<int32_t>((size_t)4 * g);
            int64_t* tab_off = m->scratch.take<int64_t>((size_t)5 * (g + 1));
            float* stats_buf = m->scratch.take<float>((size_t)g * kEmbPooled);
            WDR_REQUIRE(tab_T && tab_off && stats_buf, "embedding scratch under-reserved");
            rc = emb_forward(m, feats.p, fo, emb.p, tab_T, tab_off, stats_buf, st);
        }
        if (rc != WDR_OK) return rc;
        for (int k = 0; k < g; k++)
            WDR_CUDA_TRY(cudaMemcpyAsync(out_dev + (size_t)live[i0 + k] * m->emb_dim, emb.p + (size_t)k * m->emb_dim, sizeof(float) * m->emb_dim, cudaMemcpyDeviceToDevice, st));
        WDR_CUDA_TRY(cudaStreamSynchronize(st));
        i0 = i1;
    }
    return WDR_OK;
}

extern "C" int wdr_emb_compute_batch_i16(wdr_emb* m, const int16_t* pcm, const int64_t* seg_offset, int n_segments, float* out, int32_t* status) {
    clear_error();
    WDR_REQUIRE(m && n_segments >= 0, "bad arguments");
    if (n_segments == 0) return WDR_OK;
    WDR_REQUIRE(pcm && seg_offset && out, "null pointer");
    int rc = ensure_device(m->device);
    if (rc != WDR_OK) return rc;
    std::vector<int64_t> so(seg_offset, seg_offset + n_segments + 1);
    for (int s = 0; s < n_segments; s++) WDR_REQUIRE(so[s + 1] >= so[s], "segment offsets must ascend");
    const int64_t base = so[0], total = so[n_segments] - base;
    for (auto& v : so) v -= base;
    struct { int16_t* p; } d_pcm;
    struct { float* p; } d_out;
    if ((rc = m->io.reserve(DevArena::padded(sizeof(int16_t) * (size_t)std::max<int64_t>(total, 1)) + DevArena::padded(sizeof(float) * (size_t)n_segments * m->emb_dim))) != WDR_OK) return rc;
    d_pcm.p = m->io.take<int16_t>((size_t)std::max<int64_t>(total, 1));
    d_out.p = m->io.take<float>((size_t)n_segments * m->emb_dim);
    WDR_CUDA_TRY(cudaMemsetAsync(d_out.p, 0, sizeof(float) * (size_t)n_segments * m->emb_dim, m->stream));
    WDR_CUDA_TRY(cudaMemcpyAsync(d_pcm.p, pcm + base, sizeof(int16_t) * (size_t)total, cudaMemcpyHostToDevice, m->stream));
    rc = emb_compute_dev(m, d_pcm.p, so, d_out.p, status, m->stream);
    if (rc != WDR_OK) return rc;
    WDR_CUDA_TRY(cudaMemcpyAsync(out, d_out.p, sizeof(float) * (size_t)n_segments * m->emb_dim, cudaMemcpyDeviceToHost, m->stream));
    WDR_CUDA_TRY(cudaStreamSynchronize(m->stream));
    return WDR_OK;
}

extern "C" int wdr_emb_compute_batch_i16_dev(wdr_emb* m, const int16_t* pcm_dev, const int64_t* seg_offset_host, int n_segments, float* out_dev,
                                             int32_t* status_host, void* stream) {
    clear_error();
    WDR_REQUIRE(m && n_segments >= 0, "bad arguments");
    if (n_segments == 0) return WDR_OK;
    WDR_REQUIRE(pcm_dev && seg_offset_host && out_dev, "null pointer");
    int rc = ensure_device(m->device);
    if (rc != WDR_OK) return rc;
    std::vector<int64_t> so(seg_offset_host, seg_offset_host + n_segments + 1);
    for (int s = 0; s < n_segments; s++) WDR_REQUIRE(so[s + 1] >= so[s], "segment offsets must ascend");
    return emb_compute_dev(m, pcm_dev, so, out_dev, status_host, stream ? (cudaStream_t)stream : m->stream);
}

extern "C" int wdr_emb_compute_i16(wdr_emb* m, const int16_t* pcm, int64_t n, float* out) {
    clear_error();
    WDR_REQUIRE(m && n >= 0 && out, "bad arguments");
    if (wdr_fbank_frames((int)std::min<int64_t>(n, 1 << 30)) == 0) {
        set_error("segment of %lld samples is shorter than one 25 ms fbank frame", (long long)n);
        return WDR_ERR_TOO_SHORT;
    }
    const int64_t off[2] = {0, n};
    int32_t status = 0;
    return wdr_emb_compute_batch_i16(m, pcm, off, 1, out, &status);
}
