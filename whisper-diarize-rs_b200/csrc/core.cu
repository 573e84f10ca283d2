// core.cu — library-wide state: error text, logging hook, launch counter, device checks.
#include <stdarg.h>
#include <string.h>
#include <mutex>
#include <vector>
#include "common.cuh"

namespace wdr {

static thread_local char t_error[1024] = "";
std::atomic<uint64_t> g_launches{0};
static wdr_log_callback g_log_cb = nullptr;
static void* g_log_ud = nullptr;
static std::mutex g_log_mu;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_error, sizeof(t_error), fmt, ap);
    va_end(ap);
    log_msg(2, "%s", t_error);
}
void clear_error() { t_error[0] = 0; }

void log_msg(int level, const char* fmt, ...) {
    std::lock_guard<std::mutex> lk(g_log_mu);
    if (!g_log_cb) return;
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_log_cb(level, buf, g_log_ud);
}

// ---- DevBuf block cache ----
namespace {
struct CachedBlock { void* p; size_t bytes; int dev; };
std::mutex g_cache_mu;
std::vector<CachedBlock> g_cache_free;                 // released blocks
std::vector<CachedBlock> g_cache_live;                 // blocks handed out (to recover their size on release)
size_t g_cache_free_bytes = 0;
constexpr size_t kCacheMaxBytes = (size_t)16 << 30;    // beyond this, released blocks really go back to the driver
}  // namespace

void* devbuf_acquire(size_t bytes) {
    int dev = 0;
    cudaGetDevice(&dev);
    const size_t want = (bytes + 255) & ~(size_t)255;
    {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        int best = -1;
        for (int i = 0; i < (int)g_cache_free.size(); i++) {
            const CachedBlock& b = g_cache_free[i];
            if (b.dev == dev && b.bytes >= want && b.bytes <= 2 * want + 4096 && (best < 0 || b.bytes < g_cache_free[best].bytes)) best = i;
        }
        if (best >= 0) {
            CachedBlock b = g_cache_free[best];
            g_cache_free.erase(g_cache_free.begin() + best);
            g_cache_free_bytes -= b.bytes;
            g_cache_live.push_back(b);
            return b.p;
        }
    }
    void* p = nullptr;
    if (cudaMalloc(&p, want) != cudaSuccess) {
        devbuf_trim();  // give the cached blocks back and retry once
        cudaGetLastError();
        if (cudaMalloc(&p, want) != cudaSuccess) return nullptr;
    }
    std::lock_guard<std::mutex> lk(g_cache_mu);
    g_cache_live.push_back({p, want, dev});
    return p;
}

void devbuf_release(void* p) {
    if (!p) return;
    cudaDeviceSynchronize();  // cudaFree's implicit guarantee: nothing in flight still touches the block
    std::lock_guard<std::mutex> lk(g_cache_mu);
    for (int i = 0; i < (int)g_cache_live.size(); i++)
        if (g_cache_live[i].p == p) {
            CachedBlock b = g_cache_live[i];
            g_cache_live.erase(g_cache_live.begin() + i);
            if (g_cache_free_bytes + b.bytes > kCacheMaxBytes) { cudaFree(b.p); return; }
            g_cache_free.push_back(b);
            g_cache_free_bytes += b.bytes;
            return;
        }
    cudaFree(p);  // not ours (cannot happen through DevBuf)
}

void devbuf_trim() {
    std::vector<CachedBlock> blocks;
    {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        blocks.swap(g_cache_free);
        g_cache_free_bytes = 0;
    }
    int cur = 0;
    cudaGetDevice(&cur);
    for (auto& b : blocks) { cudaSetDevice(b.dev); cudaFree(b.p); }
    cudaSetDevice(cur);
}

int ensure_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        set_error("no CUDA device available (%s); libwdr_b200 has no CPU path", e == cudaSuccess ? "count = 0" : cudaGetErrorString(e));
        return WDR_ERR_NO_DEVICE;
    }
    if (device >= n) {
        set_error("device %d requested but only %d visible", device, n);
        return WDR_ERR_INVALID;
    }
    if (device >= 0) {
        e = cudaSetDevice(device);
        if (e != cudaSuccess) {
            set_error("cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
            return WDR_ERR_CUDA;
        }
    }
    // A non-sticky error left behind by an unrelated earlier runtime call (another library, a failed probe, an object destroyed out
    // of order) must not be reported by this call's first launch check: every compute entry point starts here, so start clean.
    cudaGetLastError();
    // Stream-ordered allocations (cudaMallocAsync in fbank / DTW helpers) come from the device's default pool, whose release
    // threshold is 0: every synchronisation hands the freed memory back to the OS and the next call maps it again — measured as
    // erratic 0.2-3 s stalls per call.  Keep freed blocks cached in the pool instead (once per device).
    {
        static std::mutex mu;
        static bool done[64] = {false};
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64) {
            std::lock_guard<std::mutex> lk(mu);
            if (!done[dev]) {
                cudaMemPool_t pool;
                if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
                    uint64_t keep = UINT64_MAX;
                    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
                }
                cudaGetLastError();
                done[dev] = true;
            }
        }
    }
    return WDR_OK;
}

}  // namespace wdr

extern "C" const char* wdr_version(void) { return "wdr_b200 0.1.0 (sm_100a)"; }
extern "C" const char* wdr_last_error(void) { return wdr::t_error; }
extern "C" int wdr_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}
extern "C" void wdr_log_set(wdr_log_callback cb, void* user_data) {
    std::lock_guard<std::mutex> lk(wdr::g_log_mu);
    wdr::g_log_cb = cb;
    wdr::g_log_ud = user_data;
}
extern "C" uint64_t wdr_launch_count(void) { return wdr::g_launches.load(); }
