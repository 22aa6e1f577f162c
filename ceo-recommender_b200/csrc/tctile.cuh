// Operand tiles of the tcgen05 tower kernels.
//
// Every matrix a tower stage multiplies (gathered inputs, activations, gradients, weights) is kept in shared memory in
// its natural row-major [R rows][C columns] fp32 shape, cut into column blocks of 32 floats: block b is R rows of 128
// bytes, 128-byte swizzled (16-byte chunk index XOR row & 7), blocks R*128 bytes apart, base 1024-byte aligned, R a
// multiple of 8.  One image serves both orientations of a tensor-core operand:
//   * K-major  (reduction runs along the columns): rows are the M/N extent; a K step of 8 floats is 32 bytes inside a
//     block row, four steps per block;
//   * MN-major (reduction runs along the rows):    columns are the M/N extent; a K step is one 8-row group (1024 B).
// fp32-class products come from two images per matrix, hi = tf32(x) and lo = x - hi, and three tensor-core passes
// a_lo.b_hi + a_hi.b_lo + a_hi.b_hi accumulated in TMEM (the towers' "precision 0"); "precision 1" runs the last one.
#pragma once
#include "tc05.cuh"

namespace cfm {

// byte offset of element (r, c) inside an image of R rows
__host__ __device__ __forceinline__ uint32_t sw128_off(int r, int c, int R) {
    return (uint32_t)((c >> 5) * R * 128 + r * 128 + ((((c & 31) >> 2) ^ (r & 7)) << 4) + ((c & 3) << 2));
}
// byte offset of the 16-byte chunk q (0..7) of row r in column block b
__host__ __device__ __forceinline__ uint32_t sw128_chunk(int r, int b, int q, int R) {
    return (uint32_t)(b * R * 128 + r * 128 + ((q ^ (r & 7)) << 4));
}

__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    uint32_t h;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
    hi = __uint_as_float(h);
    lo = x - hi;
}
__device__ __forceinline__ void split_tf32x4(const float4& x, float4& hi, float4& lo) {
    split_tf32(x.x, hi.x, lo.x);
    split_tf32(x.y, hi.y, lo.y);
    split_tf32(x.z, hi.z, lo.z);
    split_tf32(x.w, hi.w, lo.w);
}

// descriptor of K step `ks` (columns 8*ks .. 8*ks+7) of an image used K-major
__device__ __forceinline__ uint64_t tile_desc_k(uint32_t img_saddr, int R, int ks) {
    return desc_kmajor_sw128(img_saddr + (uint32_t)(ks >> 2) * (uint32_t)R * 128u + (uint32_t)(ks & 3) * 32u);
}
// descriptor of K step `ks` (rows 8*ks .. 8*ks+7) of an image used MN-major, starting at column block `b0`
__device__ __forceinline__ uint64_t tile_desc_mn(uint32_t img_saddr, int R, int ks, int b0 = 0) {
    return desc_mnmajor_sw128(img_saddr + (uint32_t)b0 * (uint32_t)R * 128u + (uint32_t)ks * 1024u, (uint32_t)R * 128u);
}

}  // namespace cfm
