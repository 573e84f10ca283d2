// cluster.cu — speaker assignment: pairwise cosine-similarity matrix on the device, pyannote-rs' online leader clustering
// (EmbeddingManager) and agglomerative (average-linkage) clustering on the device.
//
// Replaces pyannote_rs::EmbeddingManager::{new, search_speaker, get_best_speaker_match, get_all_speakers} as the crate calls
// them (reference src/transcribe.rs:342, 480-492) and provides the pairwise matrix + agglomerative clustering the
// north-star names.  Labels are a pure function of the similarity matrix S: given an identical S they are bit-exact against
// the checker (strict `>` threshold, ties toward the lowest speaker id / lowest (i, j) pair, fp32 arithmetic without FMA
// contraction in the linkage update).
#include <math.h>
#include <float.h>
#include <map>
#include <vector>
#include "common.cuh"

namespace wdr {

// S[i][j] = <e_i, e_j> / (|e_i| |e_j|)  (0 if either norm is 0), as EmbeddingManager::cosine_similarity.
// One warp per (i, j-tile of 32): lanes stride over D.
__global__ void cosine_norm_kernel(const float* __restrict__ e, int N, int D, float* __restrict__ norm) {
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= N) return;
    float s = 0.0f;
    for (int k = lane; k < D; k += 32) { const float v = e[(int64_t)i * D + k]; s = fmaf(v, v, s); }
    s = warp_sum(s);
    if (lane == 0) norm[i] = sqrtf(s);
}
__global__ void cosine_matrix_kernel(const float* __restrict__ e, const float* __restrict__ norm, int N, int D, float* __restrict__ S) {
    const int i = blockIdx.y, j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (j >= N) return;
    float s = 0.0f;
    for (int k = lane; k < D; k += 32) s = fmaf(e[(int64_t)i * D + k], e[(int64_t)j * D + k], s);
    s = warp_sum(s);
    if (lane == 0) {
        const float na = norm[i], nb = norm[j];
        S[(int64_t)i * N + j] = (na == 0.0f || nb == 0.0f) ? 0.0f : s / (na * nb);
    }
}

// Average-linkage agglomerative clustering on a similarity matrix, one CTA.  W: working copy of S (upper triangle used),
// cnt[i]: cluster sizes (0 = merged away), parent[i]: representative.  Repeats: (i, j) = argmax_{i<j alive} W[i][j] (first
// maximum in row-major order); stop when W[i][j] <= threshold; merge j into i with the Lance-Williams update
// W[i][k] = (n_i W[i][k] + n_j W[j][k]) / (n_i + n_j).
__global__ void __launch_bounds__(1024, 1)
agglomerative_kernel(float* __restrict__ W, int N, float threshold, int32_t* __restrict__ cnt, int32_t* __restrict__ parent,
                     int32_t* __restrict__ labels) {
    __shared__ unsigned long long red[32];
    __shared__ unsigned long long s_best;
    const int tid = threadIdx.x;
    for (int i = tid; i < N; i += blockDim.x) { cnt[i] = 1; parent[i] = i; }
    __syncthreads();
    const int64_t NN = (int64_t)N * N;
    for (int it = 0; it < N - 1; it++) {
        // argmax over alive pairs i < j: key = (monotone float key << 32) | (~flat index) so the FIRST maximum wins
        unsigned long long best = 0ull;
        for (int64_t f = tid; f < NN; f += blockDim.x) {
            const int i = (int)(f / N), j = (int)(f - (int64_t)i * N);
            if (i < j && cnt[i] > 0 && cnt[j] > 0) {
                const unsigned long long key = ((unsigned long long)float_to_key(W[f]) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)f);
                best = key > best ? key : best;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long u = __shfl_xor_sync(0xffffffffu, best, o);
            best = u > best ? u : best;
        }
        if ((tid & 31) == 0) red[tid >> 5] = best;
        __syncthreads();
        if (tid < 32) {
            unsigned long long t = (tid < (blockDim.x >> 5)) ? red[tid] : 0ull;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const unsigned long long u = __shfl_xor_sync(0xffffffffu, t, o);
                t = u > t ? u : t;
            }
            if (tid == 0) s_best = t;
        }
        __syncthreads();
        const unsigned long long b = s_best;
        if (b == 0ull) break;
        const float smax = key_to_float((unsigned)(b >> 32));
        if (!(smax > threshold)) break;
        const unsigned f = 0xFFFFFFFFu - (unsigned)(b & 0xFFFFFFFFull);
        const int ci = (int)(f / (unsigned)N), cj = (int)(f - (unsigned)ci * (unsigned)N);
        const float ni = (float)cnt[ci], nj = (float)cnt[cj];
        const float den = __fadd_rn(ni, nj);
        for (int k = tid; k < N; k += blockDim.x) {
            if (k == ci || k == cj || cnt[k] == 0) continue;
            const float a = W[(int64_t)min(ci, k) * N + max(ci, k)];
            const float c = W[(int64_t)min(cj, k) * N + max(cj, k)];
            W[(int64_t)min(ci, k) * N + max(ci, k)] = __fdiv_rn(__fadd_rn(__fmul_rn(ni, a), __fmul_rn(nj, c)), den);
        }
        __syncthreads();
        if (tid == 0) { cnt[ci] += cnt[cj]; cnt[cj] = 0; }
        for (int k = tid; k < N; k += blockDim.x)
            if (parent[k] == cj) parent[k] = ci;
        __syncthreads();
    }
    __syncthreads();
    // labels 1.. in order of each cluster's smallest member (parent[] always points at the smallest index of its cluster)
    if (tid == 0) {
        int next = 1;
        for (int i = 0; i < N; i++) {
            if (parent[i] == i) labels[i] = next++;
            else labels[i] = labels[parent[i]];
        }
    }
}

}  // namespace wdr

using namespace wdr;

// ---- EmbeddingManager ----------------------------------------------------------------------------------------------
struct wdr_spk {
    size_t max_speakers;
    std::map<int, std::vector<float>> speakers;  // ordered by id: ties resolve toward the lowest id
    int next_id = 1;
};

static float cosine_sim(const float* a, const float* b, int dim) {
    float dot = 0.0f, na = 0.0f, nb = 0.0f;
    for (int i = 0; i < dim; i++) { dot += a[i] * b[i]; na += a[i] * a[i]; nb += b[i] * b[i]; }
    na = sqrtf(na);
    nb = sqrtf(nb);
    if (na == 0.0f || nb == 0.0f) return 0.0f;
    return dot / (na * nb);
}

extern "C" wdr_spk* wdr_spk_init(size_t max_speakers) {
    wdr_spk* m = new wdr_spk();
    m->max_speakers = max_speakers;
    return m;
}
extern "C" void wdr_spk_free(wdr_spk* m) { delete m; }
extern "C" int wdr_spk_count(wdr_spk* m) { return m ? (int)m->speakers.size() : 0; }

extern "C" int wdr_spk_search(wdr_spk* m, const float* emb, int dim, float threshold) {
    clear_error();
    WDR_REQUIRE(m && emb && dim > 0, "bad arguments");
    int best_id = 0;
    float best = threshold;
    for (auto& kv : m->speakers) {
        if ((int)kv.second.size() != dim) { set_error("embedding dimension changed"); return WDR_ERR_INVALID; }
        const float s = cosine_sim(emb, kv.second.data(), dim);
        if (s > best) { best = s; best_id = kv.first; }
    }
    if (best_id) return best_id;
    if (m->speakers.size() < m->max_speakers) {
        const int id = m->next_id++;
        m->speakers[id] = std::vector<float>(emb, emb + dim);
        return id;
    }
    return 0;  // None: the crate renders "?" (src/transcribe.rs:493-495)
}

extern "C" int wdr_spk_best_match(wdr_spk* m, const float* emb, int dim) {
    clear_error();
    WDR_REQUIRE(m && emb && dim > 0, "bad arguments");
    if (m->speakers.empty()) { set_error("no speakers"); return WDR_ERR_INVALID; }
    int best_id = 0;
    float best = -FLT_MAX;
    for (auto& kv : m->speakers) {
        const float s = cosine_sim(emb, kv.second.data(), dim);
        if (s > best) { best = s; best_id = kv.first; }
    }
    return best_id;
}

// The crate's per-segment policy (src/transcribe.rs:480-492) over n embeddings in order: the speaker cap reached -> best match,
// else search / create.  labels[i] = id >= 1, or 0 for None ("?").  What a multi-GPU host runs on the all-gathered table
// (wdr_allgather_embeddings): O(n * speakers * D) on the host, no n x n similarity matrix.
extern "C" int wdr_spk_assign_batch(wdr_spk* m, const float* emb, int n, int dim, float threshold, int32_t* labels) {
    clear_error();
    WDR_REQUIRE(m && (emb || n == 0) && n >= 0 && dim > 0 && labels, "bad arguments");
    for (int i = 0; i < n; i++) {
        const float* e = emb + (size_t)i * dim;
        int id;
        if (m->speakers.size() == m->max_speakers && !m->speakers.empty()) id = wdr_spk_best_match(m, e, dim);
        else id = wdr_spk_search(m, e, dim, threshold);
        if (id < 0) return id;
        labels[i] = id;
    }
    return (int)m->speakers.size();
}

// ---- batch forms over a similarity matrix ----------------------------------------------------------------------------
extern "C" int wdr_cosine_matrix(const float* emb, int N, int D, float* S) {
    clear_error();
    WDR_REQUIRE(emb && S && N > 0 && D > 0, "bad arguments");
    int rc = ensure_device(-1);
    if (rc != WDR_OK) return rc;
    DevBuf<float> d_e, d_n, d_s;
    WDR_CUDA_TRY(d_e.alloc((size_t)N * D));
    WDR_CUDA_TRY(d_n.alloc(N));
    WDR_CUDA_TRY(d_s.alloc((size_t)N * N));
    WDR_CUDA_TRY(cudaMemcpy(d_e.p, emb, sizeof(float) * (size_t)N * D, cudaMemcpyHostToDevice));
    cosine_norm_kernel<<<(N + 7) / 8, 256>>>(d_e.p, N, D, d_n.p);
    WDR_LAUNCH_CHECK();
    cosine_matrix_kernel<<<dim3((N + 7) / 8, N), 256>>>(d_e.p, d_n.p, N, D, d_s.p);
    WDR_LAUNCH_CHECK();
    WDR_CUDA_TRY(cudaMemcpy(S, d_s.p, sizeof(float) * (size_t)N * N, cudaMemcpyDeviceToHost));
    return WDR_OK;
}

// The crate's policy over segments in time order (src/transcribe.rs:480-492) expressed on S: speaker k is represented by the
// FIRST segment assigned to it (pyannote-rs never updates a stored embedding).  labels: id >= 1, or 0 for "?" (None).
extern "C" int wdr_cluster_leader(const float* S, int N, float threshold, size_t max_speakers, int32_t* labels) {
    clear_error();
    WDR_REQUIRE(S && labels && N >= 0, "bad arguments");
    std::vector<int> rep;  // rep[k] = segment index of speaker k+1
    for (int i = 0; i < N; i++) {
        int best_id = 0;
        if (rep.size() == max_speakers) {  // get_best_speaker_match
            float best = -FLT_MAX;
            for (size_t k = 0; k < rep.size(); k++) {
                const float s = S[(size_t)i * N + rep[k]];
                if (s > best) { best = s; best_id = (int)k + 1; }
            }
        } else {  // search_speaker
            float best = threshold;
            for (size_t k = 0; k < rep.size(); k++) {
                const float s = S[(size_t)i * N + rep[k]];
                if (s > best) { best = s; best_id = (int)k + 1; }
            }
            if (!best_id && rep.size() < max_speakers) { rep.push_back(i); best_id = (int)rep.size(); }
        }
        labels[i] = best_id;
    }
    return (int)rep.size();
}

extern "C" int wdr_cluster_agglomerative(const float* S, int N, float threshold, int32_t* labels) {
    clear_error();
    WDR_REQUIRE(S && labels && N > 0 && N <= 46340, "bad arguments");
    int rc = ensure_device(-1);
    if (rc != WDR_OK) return rc;
    DevBuf<float> d_w;
    DevBuf<int32_t> d_cnt, d_par, d_lab;
    WDR_CUDA_TRY(d_w.alloc((size_t)N * N));
    WDR_CUDA_TRY(d_cnt.alloc(N));
    WDR_CUDA_TRY(d_par.alloc(N));
    WDR_CUDA_TRY(d_lab.alloc(N));
    WDR_CUDA_TRY(cudaMemcpy(d_w.p, S, sizeof(float) * (size_t)N * N, cudaMemcpyHostToDevice));
    agglomerative_kernel<<<1, 1024>>>(d_w.p, N, threshold, d_cnt.p, d_par.p, d_lab.p);
    WDR_LAUNCH_CHECK();
    WDR_CUDA_TRY(cudaMemcpy(labels, d_lab.p, sizeof(int32_t) * N, cudaMemcpyDeviceToHost));
    int k = 0;
    for (int i = 0; i < N; i++) k = labels[i] > k ? labels[i] : k;
    return k;
}
