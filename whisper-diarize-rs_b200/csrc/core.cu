// core.cu — library-wide state: error text, logging hook, launch counter, device checks.
#include <stdarg.h>
#include <string.h>
#include <mutex>
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
