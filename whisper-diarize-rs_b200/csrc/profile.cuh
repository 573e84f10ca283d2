// profile.cuh — optional per-kernel-class CUDA-event timing on the launching stream (bench.py's roofline source).
#pragma once
#include <cuda_runtime.h>
#include <vector>

namespace wdr {

enum KernelClass { KC_MEL = 0, KC_MEL_AUX, KC_GEMM, KC_ATTENTION, KC_LAYERNORM, KC_DECODER, KC_DTW, KC_OTHER, KC_DEC_CROSS, KC_DEC_GEMM, KC_DEC_CROSS_BATCHED, KC_COUNT };

struct Profiler {
    bool enabled = false;
    struct Rec { int kc; cudaEvent_t a, b; };
    std::vector<Rec> recs;
    std::vector<cudaEvent_t> pool;
    cudaEvent_t get() {
        if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
        cudaEvent_t e;
        cudaEventCreate(&e);
        return e;
    }
    // Sums finished records into ms[KC_COUNT] / launches[KC_COUNT] and recycles the events.  Caller synchronised the stream.
    void collect(double* ms, int* launches) {
        for (auto& r : recs) {
            float t = 0.0f;
            if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) { ms[r.kc] += t; launches[r.kc] += 1; }
            pool.push_back(r.a);
            pool.push_back(r.b);
        }
        recs.clear();
    }
    ~Profiler() {
        for (auto& r : recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
        for (auto e : pool) cudaEventDestroy(e);
    }
};

struct ProfScope {
    Profiler* p;
    cudaStream_t st;
    cudaEvent_t a = nullptr, b = nullptr;
    int kc;
    ProfScope(Profiler* prof, int kclass, cudaStream_t s) : p(prof && prof->enabled ? prof : nullptr), st(s), kc(kclass) {
        if (p) { a = p->get(); b = p->get(); cudaEventRecord(a, st); }
    }
    ~ProfScope() {
        if (p) { cudaEventRecord(b, st); p->recs.push_back({kc, a, b}); }
    }
};

}  // namespace wdr
