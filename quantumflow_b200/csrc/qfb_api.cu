// qfb_api.cu -- library bookkeeping: version, thread-local error string, launch counter, device props.
#include <stdarg.h>
#include <atomic>
#include "qfb_common.cuh"

namespace qfb {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(uint64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int sm_count_cached() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

}  // namespace qfb

extern "C" {

int qfb_version(void) { return 100; }

const char *qfb_last_error(void) { return qfb::g_err; }

uint64_t qfb_launch_count(void) { return qfb::g_launches.load(std::memory_order_relaxed); }

int qfb_device_props(int device, int *sm_count, int *cc_major, int *cc_minor, size_t *smem_optin,
                     size_t *total_mem) {
    cudaDeviceProp p;
    QFB_CUDA(cudaGetDeviceProperties(&p, device));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (smem_optin) *smem_optin = p.sharedMemPerBlockOptin;
    if (total_mem) *total_mem = p.totalGlobalMem;
    return QFB_OK;
}

}  // extern "C"
