// dist.cu — the path's ONE multi-GPU exchange behind the C ABI (SURVEY §8e, §8b "multi-GPU"): the all-gather of per-rank speaker
// embeddings ahead of global clustering, NCCL over NVLink / NVSwitch on device-resident buffers.
//
// Everything else of the path shards with no data-path collective (30 s windows, 10 s diarization windows, speech segments:
// static contiguous blocks per rank, weights replicated), so this file is all a multi-GPU host needs besides one context per GPU.
// The reference is single-GPU (`gpu_device`, reference src/engine.rs:14, src/transcribe.rs:110-112); a host that drives N GPUs
// creates one wdr_dist per rank — from N processes (wdr_dist_get_unique_id on rank 0, the id travels over the host's own channel,
// wdr_dist_init everywhere) or from N threads of one process (the same calls, or wdr_dist_init_all).
//
// NCCL is opened at run time (dlopen "libnccl.so.2", or $WDR_NCCL_LIB): the library keeps loading on a box without NCCL, where these
// entry points return WDR_ERR_UNSUPPORTED — there is no fallback transport.
#include <dlfcn.h>
#include <string.h>
#include <stdlib.h>
#include <mutex>
#include <string>
#include <vector>
#include "common.cuh"

namespace wdr {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { kNcclSuccess = 0, kNcclInt32 = 2, kNcclFloat32 = 7 };

struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int*) = nullptr;
    bool ok = false;
    std::string why;
};

static NcclApi& nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* env = getenv("WDR_NCCL_LIB");
        const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            if (!n || !n[0]) continue;
            api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (!api.lib) { api.why = "libnccl.so.2 not found (set WDR_NCCL_LIB)"; return; }
#define WDR_SYM(field, name)                                                              \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.lib, name));              \
    if (!api.field) { api.why = std::string("NCCL symbol missing: ") + name; return; }
        WDR_SYM(GetUniqueId, "ncclGetUniqueId")
        WDR_SYM(CommInitRank, "ncclCommInitRank")
        WDR_SYM(CommInitAll, "ncclCommInitAll")
        WDR_SYM(CommDestroy, "ncclCommDestroy")
        WDR_SYM(AllGather, "ncclAllGather")
        WDR_SYM(GetErrorString, "ncclGetErrorString")
        WDR_SYM(GetVersion, "ncclGetVersion")
#undef WDR_SYM
        api.ok = true;
    });
    return api;
}

#define WDR_NCCL_TRY(expr)                                                                                     \
    do {                                                                                                       \
        const int _r = (expr);                                                                                 \
        if (_r != kNcclSuccess) {                                                                              \
            ::wdr::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, nccl().GetErrorString(_r));         \
            return WDR_ERR_CUDA;                                                                               \
        }                                                                                                      \
    } while (0)

// compacts the padded all-gather result: rank r's rows [r * n_max, r * n_max + counts[r]) -> consecutive rows (rank order)
__global__ void dist_compact_kernel(const float* __restrict__ padded, const int32_t* __restrict__ counts, int n_ranks, int n_max, int D,
                                    float* __restrict__ out, int64_t out_cap_rows) {
    __shared__ int64_t s_off[65];
    if (threadIdx.x == 0) {
        int64_t o = 0;
        for (int r = 0; r < n_ranks; r++) { s_off[r] = o; o += counts[r]; }
        s_off[n_ranks] = o;
    }
    __syncthreads();
    const int64_t total = s_off[n_ranks] < out_cap_rows ? s_off[n_ranks] : out_cap_rows;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total * D; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = e / D;
        const int c = (int)(e - row * D);
        int r = 0;
        while (r + 1 < n_ranks && row >= s_off[r + 1]) r++;
        out[e] = padded[((int64_t)r * n_max + (row - s_off[r])) * D + c];
    }
}

// rows of emb scaled to unit L2 norm (a zero row stays zero): the send buffer of the all-gather when the caller asks for
// normalised embeddings, so that the cosine matrix of the gathered table is a plain GEMM E E^T
__global__ void dist_stage_kernel(const float* __restrict__ emb, int n, int D, int normalize, float* __restrict__ dst) {
    const int row = blockIdx.x;
    if (row >= n) return;
    __shared__ float red[32];
    float s = 0.0f;
    for (int c = threadIdx.x; c < D; c += blockDim.x) { const float v = emb[(int64_t)row * D + c]; s += v * v; }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0f;
        t = warp_sum(t);
        if (threadIdx.x == 0) red[0] = t;
    }
    __syncthreads();
    const float inv = (normalize && red[0] > 0.0f) ? 1.0f / sqrtf(red[0]) : 1.0f;
    for (int c = threadIdx.x; c < D; c += blockDim.x) dst[(int64_t)row * D + c] = emb[(int64_t)row * D + c] * inv;
}

}  // namespace wdr

using namespace wdr;

struct wdr_dist {
    ncclComm_t comm = nullptr;
    int n_ranks = 1, rank = 0, device = 0;
    cudaStream_t stream = nullptr;
    int32_t* counts_dev = nullptr;   // [n_ranks]
    int32_t* counts_host = nullptr;  // pinned [n_ranks]
    float* send = nullptr;           // [n_max][D] staging (grow-only)
    float* recv = nullptr;           // [n_ranks][n_max][D]
    size_t send_cap = 0, recv_cap = 0;
    float* io = nullptr;             // host-pointer entry point: local rows in / gathered rows out (grow-only)
    size_t io_cap = 0;
};

static wdr_dist* dist_finish(ncclComm_t comm, int n_ranks, int rank, int device) {
    wdr_dist* d = new wdr_dist();
    d->comm = comm; d->n_ranks = n_ranks; d->rank = rank; d->device = device;
    if (cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMalloc(&d->counts_dev, sizeof(int32_t) * n_ranks) != cudaSuccess ||
        cudaMallocHost(reinterpret_cast<void**>(&d->counts_host), sizeof(int32_t) * n_ranks) != cudaSuccess) {
        set_error("wdr_dist_init: allocation failed");
        wdr_dist_free(d);
        return nullptr;
    }
    return d;
}

extern "C" int wdr_dist_available(void) { return nccl().ok ? 1 : 0; }

extern "C" int wdr_dist_get_unique_id(uint8_t* id /* [WDR_DIST_ID_BYTES] */) {
    clear_error();
    WDR_REQUIRE(id, "bad arguments");
    if (!nccl().ok) { set_error("NCCL is not available: %s", nccl().why.c_str()); return WDR_ERR_UNSUPPORTED; }
    ncclUniqueId u;
    WDR_NCCL_TRY(nccl().GetUniqueId(&u));
    memcpy(id, u.internal, sizeof(u.internal));
    return WDR_OK;
}

extern "C" wdr_dist* wdr_dist_init(const uint8_t* id, int n_ranks, int rank, int device) {
    clear_error();
    if (n_ranks < 1 || n_ranks > 64 || rank < 0 || rank >= n_ranks || (n_ranks > 1 && !id)) { set_error("wdr_dist_init: bad arguments"); return nullptr; }
    if (ensure_device(device) != WDR_OK) return nullptr;
    if (n_ranks == 1) return dist_finish(nullptr, 1, 0, device);  // a single rank needs no communicator: the gather is a copy
    if (!nccl().ok) { set_error("NCCL is not available: %s", nccl().why.c_str()); return nullptr; }
    ncclUniqueId u;
    memcpy(u.internal, id, sizeof(u.internal));
    ncclComm_t comm = nullptr;
    const int r = nccl().CommInitRank(&comm, n_ranks, u, rank);
    if (r != kNcclSuccess) { set_error("ncclCommInitRank: %s", nccl().GetErrorString(r)); return nullptr; }
    return dist_finish(comm, n_ranks, rank, device);
}

// One process driving n_gpus devices from its own threads (the shape of a Rust host with one engine per GPU): all communicators at once.
extern "C" int wdr_dist_init_all(int n_gpus, const int* devices, wdr_dist** out) {
    clear_error();
    WDR_REQUIRE(n_gpus >= 1 && n_gpus <= 64 && out, "bad arguments");
    int rc = ensure_device(devices ? devices[0] : 0);
    if (rc != WDR_OK) return rc;
    std::vector<int> devs(n_gpus);
    for (int i = 0; i < n_gpus; i++) devs[i] = devices ? devices[i] : i;
    std::vector<ncclComm_t> comms(n_gpus, nullptr);
    if (n_gpus > 1) {
        if (!nccl().ok) { set_error("NCCL is not available: %s", nccl().why.c_str()); return WDR_ERR_UNSUPPORTED; }
        WDR_NCCL_TRY(nccl().CommInitAll(comms.data(), n_gpus, devs.data()));
    }
    for (int i = 0; i < n_gpus; i++) {
        if ((rc = ensure_device(devs[i])) != WDR_OK) return rc;
        out[i] = dist_finish(comms[i], n_gpus, i, devs[i]);
        if (!out[i]) return WDR_ERR_CUDA;
    }
    return WDR_OK;
}

extern "C" void wdr_dist_free(wdr_dist* d) {
    if (!d) return;
    cudaSetDevice(d->device);
    if (d->stream) cudaStreamSynchronize(d->stream);
    if (d->comm && nccl().ok) nccl().CommDestroy(d->comm);
    cudaFree(d->counts_dev); cudaFree(d->send); cudaFree(d->recv); cudaFree(d->io);
    if (d->counts_host) cudaFreeHost(d->counts_host);
    if (d->stream) cudaStreamDestroy(d->stream);
    delete d;
}

extern "C" int wdr_dist_size(wdr_dist* d) { return d ? d->n_ranks : 0; }
extern "C" int wdr_dist_rank(wdr_dist* d) { return d ? d->rank : -1; }
extern "C" int wdr_dist_nccl_version(void) {
    int v = 0;
    if (nccl().ok) nccl().GetVersion(&v);
    return v;
}

template <typename T>
static int grow(T** p, size_t* cap, size_t need) {
    if (need <= *cap) return WDR_OK;
    if (*p) cudaFree(*p);
    *p = nullptr; *cap = 0;
    WDR_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(p), need * sizeof(T)));
    *cap = need;
    return WDR_OK;
}

// emb_dev [n_local][D] on this rank -> out_dev [sum_r n_r][D] in rank order on EVERY rank; counts_out[r] = n_r (host).  Ranks may
// hold different (also zero) row counts: counts travel in a first 4-byte-per-rank all-gather, then every rank sends n_max rows and
// a kernel compacts the result.  normalize != 0: rows are scaled to unit L2 norm on the way into the send buffer.  Blocking: the
// row counts must reach the host before the second collective can be sized.
extern "C" int wdr_allgather_embeddings_dev(wdr_dist* d, const float* emb_dev, int n_local, int D, int normalize, float* out_dev,
                                            int64_t out_cap_rows, int32_t* counts_out, void* stream) {
    clear_error();
    WDR_REQUIRE(d && n_local >= 0 && D > 0 && (emb_dev || n_local == 0) && out_dev && out_cap_rows >= 0, "bad arguments");
    int rc = ensure_device(d->device);
    if (rc != WDR_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;  // NULL = the legacy default stream, as everywhere in CUDA: the caller's producer kernels are ordered before the gather
    const int R = d->n_ranks;
    if (R == 1) {
        WDR_REQUIRE(out_cap_rows >= n_local, "output too small");
        if (n_local) {
            dist_stage_kernel<<<n_local, 128, 0, st>>>(emb_dev, n_local, D, normalize, out_dev);
            WDR_LAUNCH_CHECK();
        }
        if (counts_out) counts_out[0] = n_local;
        WDR_CUDA_TRY(cudaStreamSynchronize(st));
        return n_local;
    }
    const int32_t mine = n_local;
    WDR_CUDA_TRY(cudaMemcpyAsync(d->counts_dev + d->rank, &mine, sizeof(int32_t), cudaMemcpyHostToDevice, st));
    WDR_NCCL_TRY(nccl().AllGather(d->counts_dev + d->rank, d->counts_dev, 1, kNcclInt32, d->comm, st));
    WDR_CUDA_TRY(cudaMemcpyAsync(d->counts_host, d->counts_dev, sizeof(int32_t) * R, cudaMemcpyDeviceToHost, st));
    WDR_CUDA_TRY(cudaStreamSynchronize(st));
    int n_max = 0;
    int64_t total = 0;
    for (int r = 0; r < R; r++) { n_max = d->counts_host[r] > n_max ? d->counts_host[r] : n_max; total += d->counts_host[r]; if (counts_out) counts_out[r] = d->counts_host[r]; }
    WDR_REQUIRE(total <= out_cap_rows, "output too small for the gathered table");
    if (total == 0) return 0;
    if ((rc = grow(&d->send, &d->send_cap, (size_t)n_max * D)) != WDR_OK) return rc;
    if ((rc = grow(&d->recv, &d->recv_cap, (size_t)R * n_max * D)) != WDR_OK) return rc;
    WDR_CUDA_TRY(cudaMemsetAsync(d->send, 0, sizeof(float) * (size_t)n_max * D, st));
    if (n_local) {
        dist_stage_kernel<<<n_local, 128, 0, st>>>(emb_dev, n_local, D, normalize, d->send);
        WDR_LAUNCH_CHECK();
    }
    WDR_NCCL_TRY(nccl().AllGather(d->send, d->recv, (size_t)n_max * D, kNcclFloat32, d->comm, st));
    const int64_t elems = total * D;
    const int blocks = (int)((elems + 255) / 256 < 148 * 8 ? (elems + 255) / 256 : 148 * 8);
    dist_compact_kernel<<<blocks, 256, 0, st>>>(d->recv, d->counts_dev, R, n_max, D, out_dev, out_cap_rows);
    WDR_LAUNCH_CHECK();
    WDR_CUDA_TRY(cudaStreamSynchronize(st));
    return (int)total;
}

// HOST pointers: one H2D of the local rows, the device path above, one D2H of the gathered table.
extern "C" int wdr_allgather_embeddings(wdr_dist* d, const float* emb, int n_local, int D, int normalize, float* out, int64_t out_cap_rows,
                                        int32_t* counts_out) {
    clear_error();
    WDR_REQUIRE(d && n_local >= 0 && D > 0 && (emb || n_local == 0) && out && out_cap_rows >= 0, "bad arguments");
    int rc = ensure_device(d->device);
    if (rc != WDR_OK) return rc;
    const size_t need = (size_t)(n_local > 0 ? n_local : 1) * D + (size_t)(out_cap_rows > 0 ? out_cap_rows : 1) * D;
    if ((rc = grow(&d->io, &d->io_cap, need)) != WDR_OK) return rc;
    float* in_dev = d->io;
    float* out_dev = d->io + (size_t)(n_local > 0 ? n_local : 1) * D;
    if (n_local) WDR_CUDA_TRY(cudaMemcpyAsync(in_dev, emb, sizeof(float) * (size_t)n_local * D, cudaMemcpyHostToDevice, d->stream));
    const int total = wdr_allgather_embeddings_dev(d, in_dev, n_local, D, normalize, out_dev, out_cap_rows, counts_out, d->stream);
    if (total < 0) return total;
    if (total) WDR_CUDA_TRY(cudaMemcpy(out, out_dev, sizeof(float) * (size_t)total * D, cudaMemcpyDeviceToHost));
    return total;
}
