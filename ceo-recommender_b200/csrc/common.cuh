// Shared device/host helpers for libcfm_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/cfm_b200.h"

namespace cfm {

// ---- host-side error plumbing -----------------------------------------------------------
void set_error(const char* fmt, ...);

#define CFM_CHECK_CUDA(expr)                                                                   \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            ::cfm::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,              \
                             cudaGetErrorString(_e));                                          \
            return CFM_ERR_CUDA;                                                               \
        }                                                                                      \
    } while (0)

#define CFM_REQUIRE(cond, code, ...)                                                           \
    do {                                                                                       \
        if (!(cond)) {                                                                         \
            ::cfm::set_error(__VA_ARGS__);                                                     \
            return (code);                                                                     \
        }                                                                                      \
    } while (0)

#define CFM_LAUNCH_CHECK()                                                                     \
    do {                                                                                       \
        ::cfm::count_launch();                                                                 \
        cudaError_t _e = cudaGetLastError();                                                   \
        if (_e != cudaSuccess) {                                                               \
            ::cfm::set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__,          \
                             cudaGetErrorString(_e));                                          \
            return CFM_ERR_CUDA;                                                               \
        }                                                                                      \
    } while (0)

int sm_count();   // cached cudaDevAttrMultiProcessorCount of the current device

// ---- launch accounting + optional per-kernel-family event timing (bench.py's roofline leg) ----
void count_launch();
enum ProfSlot {
    PROF_FWD1 = 0, PROF_FWD2, PROF_FWD3, PROF_BWD1, PROF_BWD2, PROF_BWD3, PROF_HEAD, PROF_EMB, PROF_REDUCE,
    PROF_NCE_ROWSUM, PROF_NCE_GRAD, PROF_TOPK, PROF_TOPK_POST, PROF_ADAM, PROF_SLOTS
};
struct ProfScope {   // records a cudaEvent pair around the launches in its lifetime when profiling is on
    int idx;
    cudaStream_t stream;
    ProfScope(int slot, cudaStream_t s);
    ~ProfScope();
};

// ---- device helpers ---------------------------------------------------------------------
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

// Philox4x32-10 (Salmon et al.), counter-based: the same (key, counter) always gives the same 4 words,
// so forward and backward regenerate identical dropout masks without storing them.
struct Philox4 {
    uint32_t x, y, z, w;
};
__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                           uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    return {c0, c1, c2, c3};
}

// Dropout keep decision for element (row, col) of dropout site `site` of tower `tower_id`.
// One Philox call covers 4 consecutive columns; `thresh` = p * 2^32 (keep iff word >= thresh).
struct DropCtx {
    uint32_t k0, k1;      // key   = seed
    uint32_t o0, o1;      // offset (step counter)
    uint32_t thresh;      // floor(p * 2^32)
    float inv_keep;       // 1 / (1 - p)
    uint32_t stream_id;   // tower_id * 4 + site
    bool active;
    const unsigned long long* dev_off;   // optional device-resident step counter added to the offset (CUDA graphs)
};
__host__ __device__ __forceinline__ DropCtx make_drop(double p, bool training, uint64_t seed, uint64_t offset,
                                                      int64_t tower_id, int site) {
    DropCtx d;
    d.active = training && p > 0.0;
    d.k0 = (uint32_t)seed; d.k1 = (uint32_t)(seed >> 32);
    d.o0 = (uint32_t)offset; d.o1 = (uint32_t)(offset >> 32);
    double t = p * 4294967296.0;
    d.thresh = t >= 4294967295.0 ? 0xffffffffu : (uint32_t)t;
    d.inv_keep = d.active ? (float)(1.0 / (1.0 - p)) : 1.0f;
    d.stream_id = (uint32_t)(tower_id * 4 + site);
    d.dev_off = nullptr;
    return d;
}
// fold the device-side counter into the offset (once per thread, at kernel start)
__device__ __forceinline__ DropCtx resolve_drop(DropCtx d) {
    if (d.active && d.dev_off) {
        unsigned long long o = (((unsigned long long)d.o1 << 32) | d.o0) + *d.dev_off;
        d.o0 = (uint32_t)o; d.o1 = (uint32_t)(o >> 32);
    }
    return d;
}
// words for columns [4*c4, 4*c4+3] of `row`.  Counter = (row, dropout site, column group, step offset), each in
// its own word: masks of different steps are independent draws (no word is shared between the column group and the
// step offset, so one step's mask is never a column permutation of another's); the offset's high half goes to the key.
__host__ __device__ __forceinline__ Philox4 drop_words(const DropCtx& d, long long row, int c4) {
    return philox4x32_10((uint32_t)row, (uint32_t)((unsigned long long)row >> 32) ^ (d.stream_id << 24),
                         (uint32_t)c4, d.o0, d.k0, d.k1 ^ d.o1);
}
__host__ __device__ __forceinline__ bool drop_keep(const DropCtx& d, const Philox4& w, int lane4) {
    uint32_t v = lane4 == 0 ? w.x : lane4 == 1 ? w.y : lane4 == 2 ? w.z : w.w;
    return v >= d.thresh;
}

// Column sums over the lanes of a warp by recursive halving: every step a lane keeps one half of its columns and
// sends the other half to its partner (xor mask), so 32 lanes x NC columns cost NC - NC/32 shuffles instead of
// 5 * NC.  After masks 16..1 lane l holds the sums of columns (NC/32)*l ... ; with masks 8..1 (16-lane groups, the
// M = 64 accumulator layout) lane l < 16 holds columns (NC/16)*l ... of its group.
template <int N, int MASK>
struct Halve {
    __device__ static __forceinline__ void run(float (&x)[N], int lane) {
        const bool up = (lane & MASK) != 0;
#pragma unroll
        for (int i = 0; i < N / 2; ++i) {
            const float keep = up ? x[i + N / 2] : x[i];
            const float send = up ? x[i] : x[i + N / 2];
            x[i] = keep + __shfl_xor_sync(FULL, send, MASK);
        }
        if constexpr (MASK > 1) Halve<N / 2, MASK / 2>::run(reinterpret_cast<float (&)[N / 2]>(x), lane);
    }
};

__host__ __device__ __forceinline__ int pad_ld(int k) {   // leading dim: multiple of 4 floats, (ld/4) odd
    int ld = (k + 3) & ~3;
    if (((ld >> 2) & 1) == 0) ld += 4;
    return ld;
}
__host__ __device__ __forceinline__ int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace cfm
