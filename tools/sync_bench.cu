// sync_bench.cu — what does a dependency edge cost on this GPU?  (measurement helper, not part of the product)
//   (1) a CUDA graph of N dependent trivial kernels, plain edges          -> us per kernel boundary
//   (2) the same with programmatic dependent launch (griddepcontrol)       -> us per boundary
//   (3) the same, every CTA holding 200 KB of shared memory (two such CTAs cannot share an SM, like the decode GEMMs)
//   (4) ONE persistent kernel (one CTA per SM) with N grid-wide barriers (release add + acquire poll in L2) -> us per barrier
// Every "phase" does the same token amount of dependent work: each CTA reads a value another CTA wrote in the previous phase.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/sync_bench tools/sync_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__global__ void phase_kernel(const float* __restrict__ in, float* __restrict__ out, int pdl) {
    extern __shared__ float sm[];
    if (pdl) {
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        asm volatile("griddepcontrol.wait;" ::: "memory");
    }
    const int peer = (blockIdx.x + 1) % gridDim.x;
    const float v = __ldcg(in + peer * blockDim.x + threadIdx.x);
    out[blockIdx.x * blockDim.x + threadIdx.x] = v + 1.0f;
}

__global__ void small_kernel(const float* __restrict__ in, float* __restrict__ out, int pdl) {
    if (pdl) {
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        asm volatile("griddepcontrol.wait;" ::: "memory");
    }
    const int peer = (blockIdx.x + 1) % gridDim.x;
    const float v = __ldcg(in + peer * blockDim.x + threadIdx.x);
    out[blockIdx.x * blockDim.x + threadIdx.x] = v + 1.0f;
}

__device__ __forceinline__ void grid_barrier(unsigned* ctr, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
        unsigned v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
        } while (v < target);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(384, 1) persistent_kernel(float* a, float* b, unsigned* ctr, int phases) {
    extern __shared__ float sm[];
    const int peer = (blockIdx.x + 1) % gridDim.x;
    float* in = a;
    float* out = b;
    for (int ph = 0; ph < phases; ph++) {
        const float v = __ldcg(in + peer * blockDim.x + threadIdx.x);
        out[blockIdx.x * blockDim.x + threadIdx.x] = v + 1.0f;
        grid_barrier(ctr, (unsigned)(ph + 1) * gridDim.x);
        float* t = in; in = out; out = t;
    }
}

// alternating: even kernels hold `smem` bytes (phase_kernel), odd kernels are small_kernel with no shared memory
static float run_graph(int n, int pdl, size_t smem, int grid, int block, float* a, float* b, cudaStream_t st, bool alternate = false) {
    CK(cudaFuncSetAttribute(phase_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    cudaGraph_t g;
    cudaGraphExec_t ge;
    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    for (int i = 0; i < n; i++) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at;
        cfg.numAttrs = pdl ? 1 : 0;
        const float* in = (i & 1) ? b : a;
        float* out = (i & 1) ? a : b;
        if (alternate && (i & 1)) {
            cfg.dynamicSmemBytes = 0;
            CK(cudaLaunchKernelEx(&cfg, small_kernel, in, out, pdl));
        } else
        CK(cudaLaunchKernelEx(&cfg, phase_kernel, in, out, pdl));
    }
    CK(cudaStreamEndCapture(st, &g));
    CK(cudaGraphInstantiate(&ge, g, 0));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 3; w++) CK(cudaGraphLaunch(ge, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaEventRecord(e0, st));
    const int reps = 10;
    for (int r = 0; r < reps; r++) CK(cudaGraphLaunch(ge, st));
    CK(cudaEventRecord(e1, st));
    CK(cudaStreamSynchronize(st));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    CK(cudaGraphExecDestroy(ge)); CK(cudaGraphDestroy(g));
    return ms * 1000.0f / (reps * n);
}

int main() {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    cudaStream_t st;
    CK(cudaStreamCreate(&st));
    const int block = 384, n = 320;
    float *a, *b;
    unsigned* ctr;
    CK(cudaMalloc(&a, sms * block * 4)); CK(cudaMalloc(&b, sms * block * 4)); CK(cudaMalloc(&ctr, 4));
    CK(cudaMemset(a, 0, sms * block * 4)); CK(cudaMemset(b, 0, sms * block * 4));
    printf("{\"sms\": %d", sms);
    printf(", \"graph_plain_us\": %.3f", run_graph(n, 0, 0, sms, block, a, b, st));
    printf(", \"graph_pdl_us\": %.3f", run_graph(n, 1, 0, sms, block, a, b, st));
    printf(", \"graph_plain_200k_us\": %.3f", run_graph(n, 0, 200 * 1024, sms, block, a, b, st));
    printf(", \"graph_pdl_200k_us\": %.3f", run_graph(n, 1, 200 * 1024, sms, block, a, b, st));
    printf(", \"graph_pdl_120cta_us\": %.3f", run_graph(n, 1, 0, 120, 320, a, b, st));
    printf(", \"graph_pdl_alt_200k_0k_us\": %.3f", run_graph(n, 1, 200 * 1024, sms, block, a, b, st, true));
    printf(", \"graph_plain_alt_200k_0k_us\": %.3f", run_graph(n, 0, 200 * 1024, sms, block, a, b, st, true));
    CK(cudaFuncSetAttribute(small_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    printf(", \"graph_pdl_alt_200k_0k_maxshared_us\": %.3f", run_graph(n, 1, 200 * 1024, sms, block, a, b, st, true));
    printf(", \"graph_plain_alt_200k_0k_maxshared_us\": %.3f", run_graph(n, 0, 200 * 1024, sms, block, a, b, st, true));
    CK(cudaFuncSetAttribute(persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (size_t smem : {(size_t)0, (size_t)200 * 1024}) {
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        float best = 1e9f;
        for (int r = 0; r < 5; r++) {
            CK(cudaMemsetAsync(ctr, 0, 4, st));
            CK(cudaEventRecord(e0, st));
            persistent_kernel<<<sms, block, smem, st>>>(a, b, ctr, n);
            CK(cudaEventRecord(e1, st));
            CK(cudaStreamSynchronize(st));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (ms < best) best = ms;
        }
        printf(", \"persistent_barrier%s_us\": %.3f", smem ? "_200k" : "", best * 1000.0f / n);
    }
    printf("}\n");
    return 0;
}
