// Library-level plumbing of libcfm_b200: error strings, device queries.
#include "common.cuh"
#include <string.h>

namespace cfm {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count() {
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

}  // namespace cfm

extern "C" int cfm_abi_version(void) { return CFM_ABI_VERSION; }

extern "C" const char* cfm_last_error(void) { return cfm::g_err; }

extern "C" int cfm_device_info(int64_t* sm, int64_t* cc_major, int64_t* cc_minor, int64_t* tower_ctas) {
    int dev = 0, maj = 0, min = 0;
    CFM_CHECK_CUDA(cudaGetDevice(&dev));
    CFM_CHECK_CUDA(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev));
    CFM_CHECK_CUDA(cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev));
    if (sm) *sm = cfm::sm_count();
    if (cc_major) *cc_major = maj;
    if (cc_minor) *cc_minor = min;
    if (tower_ctas) *tower_ctas = cfm::sm_count();
    return CFM_OK;
}
