// Inline-PTX wrappers for the Blackwell (sm_100a) asynchronous machinery shared by the tensor-core kernels:
// mbarriers, TMA, tcgen05 (MMA issue / commit / TMEM alloc + loads), shared-memory matrix descriptors.
#pragma once
#include "common.cuh"
#include <cuda.h>

namespace cfm {

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// true in exactly one lane of the (converged) warp; code under `if (elect_one())` in otherwise warp-uniform control flow
// keeps its operands in uniform registers, which is what lets tcgen05.mma issue back to back
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem], bf16 inputs, fp32 accumulate; issued by ONE thread
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
// same, TF32 operands (fp32 bit patterns in shared memory, top 19 bits used), K = 8 per instruction
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
// instruction descriptor for kind::tf32: fp32 accumulate, TF32 A/B (=2 @7, @10), major bits @15/@16, N>>3 @17, M>>4 @24
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N, bool a_mn_major, bool b_mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// mbarrier arrives once every MMA issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 consecutive fp32 columns of this thread's TMEM lane (WAIT = false: the caller issues tmem_ld_wait() later)
template <bool WAIT = true>
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    if (WAIT) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// 16 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// 64 consecutive fp32 columns of this thread's TMEM lane (one instruction, one wait)
template <bool WAIT = true>
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, float (&v)[64]) {
    uint32_t r[64];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
        "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
        "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]),
          "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]),
          "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]),
          "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]),
          "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
        : "r"(taddr)
        : "memory");
    if (WAIT) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 64; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// Pins 32 values behind the preceding volatile statements (a tmem_ld_wait): arithmetic on the results of an
// un-waited TMEM load must not be scheduled above the wait.
__device__ __forceinline__ void reg_fence32(float (&v)[32]) {
    asm volatile("" : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]),
                      "+f"(v[8]), "+f"(v[9]), "+f"(v[10]), "+f"(v[11]), "+f"(v[12]), "+f"(v[13]), "+f"(v[14]), "+f"(v[15]),
                      "+f"(v[16]), "+f"(v[17]), "+f"(v[18]), "+f"(v[19]), "+f"(v[20]), "+f"(v[21]), "+f"(v[22]), "+f"(v[23]),
                      "+f"(v[24]), "+f"(v[25]), "+f"(v[26]), "+f"(v[27]), "+f"(v[28]), "+f"(v[29]), "+f"(v[30]), "+f"(v[31])
                 :: "memory");
}
// Streams the 128 fp32 columns of this thread's TMEM lane through registers in four 32-column pieces: piece k + 1 is
// in flight while f(k, piece) runs; released() is called as soon as the last piece has landed (the accumulator
// buffer may be handed back to the MMA issuer while the last piece is still being worked on).
template <class F, class R>
__device__ __forceinline__ void tmem_stream128(uint32_t taddr, F&& f, R&& released) {
    float va[32], vb[32];
    tmem_ld32<false>(taddr, va);
    tmem_ld_wait(); reg_fence32(va);
    tmem_ld32<false>(taddr + 32, vb);
    f(0, va);
    tmem_ld_wait(); reg_fence32(vb);
    tmem_ld32<false>(taddr + 64, va);
    f(1, vb);
    tmem_ld_wait(); reg_fence32(va);
    tmem_ld32<false>(taddr + 96, vb);
    f(2, va);
    tmem_ld_wait(); reg_fence32(vb);
    released();
    f(3, vb);
}
// 2^x on the SFU (MUFU.EX2), flush-to-zero: one instruction, no denormal fix-up
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// 2^x on the FMA/ALU pipes for x in [-125, 0]: round-to-nearest split x = n + f (f in [-0.5, 0.5]) through the
// 1.5*2^23 magic constant, degree-4 minimax polynomial for 2^f (relative error 2.7e-6), n added into the exponent
// field.  Used for a fixed quarter of the exponentials of the similarity epilogues, which are otherwise bound by the
// 16/clk/SM MUFU pipe.
__device__ __forceinline__ float ex2_poly(float x) {
    const float t = x + 12582912.f;
    const float f = x - (t - 12582912.f);
    float p = fmaf(f, 0.00957009382545948f, 0.05591785907745361f);
    p = fmaf(p, f, 0.240247443318367f);
    p = fmaf(p, f, 0.6931217908859253f);
    p = fmaf(p, f, 0.9999992847442627f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
// 16 exponentials e[i] = 2^(v[i] * c1 - c2), written phase by phase so that the instruction stream carries the
// parallelism itself (ptxas keeps source order under register pressure: an element-at-a-time loop becomes one long
// dependent chain).  POLY: elements 3, 7, 11, 15 take the polynomial, four chains interleaved with the twelve MUFUs.
template <bool POLY>
__device__ __forceinline__ void exp16(const float* v, float c1, float c2, float (&e)[16]) {
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = fmaf(v[i], c1, -c2);
    if (!POLY) {
#pragma unroll
        for (int i = 0; i < 16; ++i) e[i] = ex2_approx(x[i]);
        return;
    }
    float t[4], f[4], p[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) t[k] = x[4 * k + 3] + 12582912.f;
    e[0] = ex2_approx(x[0]); e[1] = ex2_approx(x[1]);
#pragma unroll
    for (int k = 0; k < 4; ++k) f[k] = t[k] - 12582912.f;
    e[2] = ex2_approx(x[2]); e[4] = ex2_approx(x[4]);
#pragma unroll
    for (int k = 0; k < 4; ++k) f[k] = x[4 * k + 3] - f[k];
    e[5] = ex2_approx(x[5]); e[6] = ex2_approx(x[6]);
#pragma unroll
    for (int k = 0; k < 4; ++k) p[k] = fmaf(f[k], 0.00957009382545948f, 0.05591785907745361f);
    e[8] = ex2_approx(x[8]); e[9] = ex2_approx(x[9]);
#pragma unroll
    for (int k = 0; k < 4; ++k) p[k] = fmaf(p[k], f[k], 0.240247443318367f);
    e[10] = ex2_approx(x[10]); e[12] = ex2_approx(x[12]);
#pragma unroll
    for (int k = 0; k < 4; ++k) p[k] = fmaf(p[k], f[k], 0.6931217908859253f);
    e[13] = ex2_approx(x[13]); e[14] = ex2_approx(x[14]);
#pragma unroll
    for (int k = 0; k < 4; ++k) p[k] = fmaf(p[k], f[k], 0.9999992847442627f);
#pragma unroll
    for (int k = 0; k < 4; ++k) e[4 * k + 3] = __int_as_float(__float_as_int(p[k]) + (__float_as_int(t[k]) << 23));
}

// Shared-memory matrix descriptors (sm_100 format: version 1 at bit 46, layout type at [61,64), SWIZZLE_128B = 2).
// K-major operand stored as [rows][64 bf16] 128-byte rows, 8-row groups 1024 B apart (SBO); LBO unused.
__device__ __forceinline__ uint64_t desc_kmajor_sw128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// MN-major operand: 64 contiguous bf16 along MN per 128-byte row, rows = K; 8-row (K) groups 1024 B apart (SBO),
// next 64 elements along MN `lbo_bytes` away (LBO).
__device__ __forceinline__ uint64_t desc_mnmajor_sw128(uint32_t saddr, uint32_t lbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor: fp32 accumulate (c_format=1 @4), bf16 A/B (=1 @7, @10), b_major @16, N>>3 @17, M>>4 @24
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, bool b_mn_major, bool f16 = false) {
    const uint32_t fmt = f16 ? 0u : 1u;        // kind::f16 operand format: 0 = fp16, 1 = bf16
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((b_mn_major ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}


}  // namespace cfm
