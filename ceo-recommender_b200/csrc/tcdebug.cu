// Self-test of the tcgen05 kind::tf32 building blocks the tower kernels (tower_tc.cu) rest on: one CTA multiplies two
// small fp32 matrices that its threads have split into (hi, lo) TF32 halves and laid out in shared memory as
// 128-byte-swizzled column blocks, in every operand orientation the towers use, and dumps the raw TMEM accumulator.
// tests/test_gpu_tc_blocks.py compares the dump with a float64 product: it pins the descriptor conventions (K-major
// vs MN-major, M = 64 lane mapping) and the accuracy of the three-pass error-compensated product.
#include "common.cuh"
#include "tc05.cuh"
#include "tctile.cuh"

namespace cfm {

// mode 0: D[M,N] = A[M,K] . B[N,K]^T      (A, B K-major)        forward Linear
// mode 1: D[M,N] = A[K,M]^T . B[K,N]      (A, B MN-major)       weight gradient
// mode 2: D[M,N] = A[M,K] . B[K,N]        (A K-major, B MN-major) input gradient
__global__ void __launch_bounds__(128, 1) tc_selftest_kernel(const float* __restrict__ A, const float* __restrict__ Bm,
                                                             float* __restrict__ out, int mode, int M, int N, int K,
                                                             int passes) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic on the __shared__ array: accesses compile to LDS/STS (a uintptr_t round trip makes them generic LD/ST)
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int Ra = mode == 1 ? K : M, Ca = mode == 1 ? M : K;       // stored shape of A
    const int Rb = mode == 0 ? N : K, Cb = mode == 0 ? K : N;       // stored shape of B
    const int Rap = (Ra + 7) & ~7, Rbp = (Rb + 7) & ~7;
    const int nba = (Ca + 31) >> 5, nbb = (Cb + 31) >> 5;
    const uint32_t a_img = Rap * 128 * nba, b_img = Rbp * 128 * nbb;
    uint8_t *a_hi = sm, *a_lo = sm + a_img, *b_hi = sm + 2 * a_img, *b_lo = sm + 2 * a_img + b_img;
    for (uint32_t i = tid; i < (2 * a_img + 2 * b_img) / 16; i += 128) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    for (int i = tid; i < Ra * Ca; i += 128) {
        const int r = i / Ca, c = i - r * Ca;
        float hi, lo;
        split_tf32(A[i], hi, lo);
        *reinterpret_cast<float*>(a_hi + sw128_off(r, c, Rap)) = hi;
        *reinterpret_cast<float*>(a_lo + sw128_off(r, c, Rap)) = lo;
    }
    for (int i = tid; i < Rb * Cb; i += 128) {
        const int r = i / Cb, c = i - r * Cb;
        float hi, lo;
        split_tf32(Bm[i], hi, lo);
        *reinterpret_cast<float*>(b_hi + sw128_off(r, c, Rbp)) = hi;
        *reinterpret_cast<float*>(b_lo + sw128_off(r, c, Rbp)) = lo;
    }
    if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    uint32_t cols = 32;
    while (cols < (uint32_t)N) cols <<= 1;
    if (warp == 0) tmem_alloc(&tmem_slot, cols);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (tid == 0) {
        const uint32_t idesc = make_idesc_tf32(M, N, mode == 1, mode != 0);
        const int ksteps = (K + 7) >> 3;
        for (int p = 0; p < passes; ++p) {
            // small terms first: a_lo.b_hi, a_hi.b_lo, then a_hi.b_hi  (single pass: a_hi.b_hi only)
            const int which = passes == 1 ? 2 : p;
            const uint32_t ab = smem_u32(which == 0 ? a_lo : a_hi), bb = smem_u32(which == 1 ? b_lo : b_hi);
            for (int ks = 0; ks < ksteps; ++ks) {
                const uint64_t da = mode == 1 ? tile_desc_mn(ab, Rap, ks) : tile_desc_k(ab, Rap, ks);
                const uint64_t db = mode == 0 ? tile_desc_k(bb, Rbp, ks) : tile_desc_mn(bb, Rbp, ks);
                umma_tf32(tmem, da, db, idesc, p > 0 || ks > 0);
            }
        }
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < (int)cols; c0 += 32) {
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(32 * warp) << 16) + c0, v);
        for (int i = 0; i < 32; ++i) out[(size_t)tid * cols + c0 + i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, cols);
}

}  // namespace cfm

using namespace cfm;

extern "C" int cfm_tc_selftest(const float* A, const float* B, float* out, int64_t mode, int64_t M, int64_t N,
                               int64_t K, int64_t passes, void* stream_) {
    CFM_REQUIRE(A && B && out, CFM_ERR_INVALID, "null buffer");
    CFM_REQUIRE(mode >= 0 && mode <= 2 && (M == 64 || M == 128) && N >= 8 && N <= 256 && (N % (M == 128 ? 16 : 8)) == 0 &&
                    K >= 8 && (K % 8) == 0 && (passes == 1 || passes == 3),
                CFM_ERR_INVALID, "bad self-test shape");
    const int64_t Ra = mode == 1 ? K : M, Ca = mode == 1 ? M : K, Rb = mode == 0 ? N : K, Cb = mode == 0 ? K : N;
    const size_t smem = 2 * (size_t)((Ra + 7) & ~7) * 128 * ((Ca + 31) / 32) + 2 * (size_t)((Rb + 7) & ~7) * 128 * ((Cb + 31) / 32) + 1024;
    CFM_REQUIRE(smem <= 227 * 1024, CFM_ERR_UNSUPPORTED, "self-test operands need %zu B of shared memory", smem);
    CFM_CHECK_CUDA(cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream_>>>(A, B, out, (int)mode, (int)M, (int)N, (int)K, (int)passes);
    CFM_LAUNCH_CHECK();
    return CFM_OK;
}

// Timing probe: `reps` back-to-back tcgen05.mma kind::tf32 (M x N x 8 each, operands = whatever is in shared memory)
// rotating over `nacc` TMEM accumulators; out[0] = cycles from first issue to completion (one thread's clock64).
namespace cfm {
__global__ void __launch_bounds__(128, 1) tc_mma_probe(long long* out, int M, int N, int reps, int nacc) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic on the __shared__ array: accesses compile to LDS/STS (a uintptr_t round trip makes them generic LD/ST)
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 48 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (warp == 0) {
        // whole warp walks the loop (uniform control flow); one elected lane issues
        const uint32_t idesc = make_idesc_tf32(M, N, false, false);
        const uint32_t a = smem_u32(sm), b = smem_u32(sm) + 16384;
        const long long t0 = clock64();
        int acc = 0;
        for (int r = 0; r < reps; ++r) {
            if (elect_one())
                umma_tf32(tmem + (uint32_t)acc * (uint32_t)N, desc_kmajor_sw128(a + (r & 3) * 32), desc_kmajor_sw128(b + (r & 3) * 32), idesc, r >= nacc);
            __syncwarp();
            if (++acc == nacc) acc = 0;
        }
        const long long t1 = clock64();
        if (elect_one()) umma_commit(&bar);
        __syncwarp();
        mbar_wait(&bar, 0);
        const long long t2 = clock64();
        if (tid == 0) { out[0] = t2 - t0; out[1] = t1 - t0; }
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}
}  // namespace cfm

extern "C" int cfm_tc_mma_probe(long long* out, int64_t M, int64_t N, int64_t reps, int64_t nacc, void* stream_) {
    CFM_REQUIRE(out && (M == 64 || M == 128) && N >= 8 && N <= 256 && nacc >= 1 && nacc * N <= 512, CFM_ERR_INVALID, "bad probe shape");
    CFM_CHECK_CUDA(cudaFuncSetAttribute(cfm::tc_mma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 50 * 1024));
    cfm::tc_mma_probe<<<1, 128, 50 * 1024, (cudaStream_t)stream_>>>(out, (int)M, (int)N, (int)reps, (int)nacc);
    CFM_LAUNCH_CHECK();
    return CFM_OK;
}
