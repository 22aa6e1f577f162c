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

// ---- launch counter and event-pair profiling ----
static long long g_launches = 0;
void count_launch() { ++g_launches; }

constexpr int PROF_MAX = 8192;
static bool g_prof_on = false;
static cudaEvent_t g_ev[PROF_MAX][2];
static int g_ev_slot[PROF_MAX];
static int g_ev_created = 0, g_ev_used = 0;

ProfScope::ProfScope(int slot, cudaStream_t s) : idx(-1), stream(s) {
    if (!g_prof_on || g_ev_used >= PROF_MAX) return;
    if (g_ev_used >= g_ev_created) {
        if (cudaEventCreate(&g_ev[g_ev_created][0]) != cudaSuccess) return;
        if (cudaEventCreate(&g_ev[g_ev_created][1]) != cudaSuccess) return;
        ++g_ev_created;
    }
    idx = g_ev_used++;
    g_ev_slot[idx] = slot;
    cudaEventRecord(g_ev[idx][0], stream);
}
ProfScope::~ProfScope() {
    if (idx >= 0) cudaEventRecord(g_ev[idx][1], stream);
}

}  // namespace cfm

extern "C" int64_t cfm_launch_count(int64_t reset) {
    long long v = cfm::g_launches;
    if (reset) cfm::g_launches = 0;
    return v;
}

extern "C" int cfm_profile_enable(int64_t on) {
    cfm::g_prof_on = on != 0;
    cfm::g_ev_used = 0;
    return CFM_OK;
}

extern "C" int cfm_profile_read(double* ms, int64_t* counts, int64_t n_slots) {
    using namespace cfm;
    CFM_REQUIRE(ms && counts && n_slots >= PROF_SLOTS, CFM_ERR_INVALID, "need %d profile slots", (int)PROF_SLOTS);
    for (int i = 0; i < n_slots; ++i) { ms[i] = 0.0; counts[i] = 0; }
    CFM_CHECK_CUDA(cudaDeviceSynchronize());
    for (int i = 0; i < g_ev_used; ++i) {
        float t = 0.f;
        CFM_CHECK_CUDA(cudaEventElapsedTime(&t, g_ev[i][0], g_ev[i][1]));
        ms[g_ev_slot[i]] += t;
        counts[g_ev_slot[i]] += 1;
    }
    g_ev_used = 0;
    return CFM_OK;
}

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
    if (tower_ctas) *tower_ctas = 2 * cfm::sm_count();   // maximum persistent CTAs per tower (32-row tile stages)
    return CFM_OK;
}
