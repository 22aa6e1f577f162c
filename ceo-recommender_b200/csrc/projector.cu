// Projection heads of the contrastive model: p = normalize(W2 . relu(W1 . x + b1) + b2)
//
// replaces ceo_firm_matching/contrastive.py:41-50 (the two nn.Sequential(Linear, ReLU, Linear) heads), :96-97
// (F.normalize of their outputs) and the autograd graph torch builds for them (:244-260 `loss.backward()`).
//
// The heads are tiny (60 -> 60 -> 30 per side, 0.7 GFLOP forward at B = 65 536) and sit between two fp32-exact
// neighbours (cosine head, InfoNCE packing), so they run as plain fp32 FMA register-tiled GEMMs: one CTA of 256
// threads per 64-row tile, both weight matrices resident in shared memory, thread (ty, tx) owns rows 4ty..4ty+3 and
// columns 4tx..4tx+3 of every 64 x 64 product.  Both heads of the model share one launch (blockIdx.y).  The backward
// keeps the CTA's weight-gradient blocks in registers over all of its tiles, writes one partial per CTA and a second
// tiny kernel adds the partials in CTA order: bitwise reproducible, no atomics.
#include "common.cuh"
#include <algorithm>

namespace cfm {

constexpr int PJ_NT = 256;
constexpr int PJ_M = 64;             // rows per tile
constexpr int PJ_D = 64;             // widest supported layer
constexpr int PJ_LD = PJ_D + 4;      // row pitch of the activation tiles (float4-aligned, bank-shifted by 4 per row)
constexpr int PJ_MAX_HEADS = 4;

struct PjHead {
    int d_in, d_hid, d_out;
    const float *w1, *b1, *w2, *b2, *x;
    float *hid, *raw, *out;
    const float* g_out;
    float *dx, *dw1, *db1, *dw2, *db2, *scratch;
};
struct PjArgs {
    PjHead h[PJ_MAX_HEADS];
    long long B;
    float eps;
};

__host__ __device__ inline int pj_partial_floats(int d_in, int d_hid, int d_out) {
    return d_hid * d_in + d_hid + d_out * d_hid + d_out;
}

// rows [row0, row0 + 64) of a row-major [B, d] matrix -> tile [64][PJ_LD], zero outside (coalesced scalar loads:
// the rows of a tile are contiguous in memory)
__device__ __forceinline__ void pj_load_tile(float* tile, const float* src, long long row0, long long B, int d, int tid) {
    const int rows = (int)min((long long)PJ_M, B - row0);
    const float* base = src + (size_t)row0 * d;
    for (int e = tid; e < PJ_M * PJ_D; e += PJ_NT) {
        const int r = e >> 6, c = e & 63;
        tile[r * PJ_LD + c] = 0.f;
    }
    __syncthreads();
    const int n = rows * d;
    for (int e = tid; e < n; e += PJ_NT) {
        const int r = e / d, c = e - r * d;
        tile[r * PJ_LD + c] = __ldg(base + e);
    }
}
// weight [n_out, n_in] (torch layout) -> smem [64][64]: transposed (dst[k][j] = w[j][k]) or as is, zero padded
__device__ __forceinline__ void pj_load_weight(float* dst, const float* w, int n_out, int n_in, bool transpose, int tid) {
    for (int e = tid; e < PJ_D * PJ_D; e += PJ_NT) dst[e] = 0.f;
    __syncthreads();
    for (int e = tid; e < n_out * n_in; e += PJ_NT) {
        const int j = e / n_in, k = e - j * n_in;
        const float v = __ldg(w + e);
        if (transpose) dst[k * PJ_D + j] = v; else dst[j * PJ_D + k] = v;
    }
}
// acc[i][c] += sum_k A[4ty + i][k] * Bm[k][4tx + c]   (A: tile with pitch PJ_LD, Bm: [K][64])
__device__ __forceinline__ void pj_gemm_rows(float (&acc)[4][4], const float* A, const float* Bm, int K, int ty, int tx) {
    const float* a0 = A + (4 * ty) * PJ_LD;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
        const float4 b = *reinterpret_cast<const float4*>(Bm + k * PJ_D + 4 * tx);
        const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float a = a0[i * PJ_LD + k];
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[i][c] = fmaf(a, bb[c], acc[i][c]);
        }
    }
}
// acc[i][c] += sum_r P[r][4ty + i] * Q[r][4tx + c]   (reduction over the tile's rows: weight gradients)
__device__ __forceinline__ void pj_gemm_cols(float (&acc)[4][4], float (&colsum)[4], const float* P, const float* Q, int ty, int tx) {
#pragma unroll 4
    for (int r = 0; r < PJ_M; ++r) {
        const float4 p = *reinterpret_cast<const float4*>(P + r * PJ_LD + 4 * ty);
        const float4 q = *reinterpret_cast<const float4*>(Q + r * PJ_LD + 4 * tx);
        const float pp[4] = {p.x, p.y, p.z, p.w}, qq[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            colsum[i] += pp[i];
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[i][c] = fmaf(pp[i], qq[c], acc[i][c]);
        }
    }
}
__device__ __forceinline__ float pj_sum16(float v) {      // over the 16 lanes that share a row group
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ void pj_store_rows(float* dst, const float (&v)[4][4], long long row0, long long B, int d, int ty, int tx) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long r = row0 + 4 * ty + i;
        if (r < B) {
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (4 * tx + c < d) dst[(size_t)r * d + 4 * tx + c] = v[i][c];
        }
    }
}

__global__ void __launch_bounds__(PJ_NT) projector_fwd_kernel(const __grid_constant__ PjArgs a) {
    extern __shared__ __align__(16) float pj_sm[];
    const PjHead& H = a.h[blockIdx.y];
    float* w1t = pj_sm;                       // [d_in][64]   w1t[k][j] = w1[j][k]
    float* w2t = w1t + PJ_D * PJ_D;           // [d_hid][64]
    float* xs = w2t + PJ_D * PJ_D;            // [64][PJ_LD]
    float* hs = xs + PJ_M * PJ_LD;
    float* bs = hs + PJ_M * PJ_LD;            // b1 [64] | b2 [64]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    pj_load_weight(w1t, H.w1, H.d_hid, H.d_in, true, tid);
    pj_load_weight(w2t, H.w2, H.d_out, H.d_hid, true, tid);
    if (tid < 64) bs[tid] = tid < H.d_hid ? H.b1[tid] : 0.f;
    else if (tid < 128) bs[tid] = tid - 64 < H.d_out ? H.b2[tid - 64] : 0.f;
    const long long ntiles = (a.B + PJ_M - 1) / PJ_M;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const long long row0 = t * PJ_M;
        __syncthreads();                      // previous tile's readers of xs / hs are done
        pj_load_tile(xs, H.x, row0, a.B, H.d_in, tid);
        __syncthreads();
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[i][c] = 0.f;
        pj_gemm_rows(acc, xs, w1t, H.d_in, ty, tx);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                acc[i][c] = fmaxf(acc[i][c] + bs[4 * tx + c], 0.f);
                hs[(4 * ty + i) * PJ_LD + 4 * tx + c] = acc[i][c];
            }
        pj_store_rows(H.hid, acc, row0, a.B, H.d_hid, ty, tx);
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[i][c] = 0.f;
        pj_gemm_rows(acc, hs, w2t, H.d_hid, ty, tx);
        float o[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float ss = 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                acc[i][c] += bs[64 + 4 * tx + c];       // padded columns: zero weights + zero bias = 0
                ss = fmaf(acc[i][c], acc[i][c], ss);
            }
            ss = pj_sum16(ss);
            const float inv = 1.f / fmaxf(sqrtf(ss), a.eps);      // F.normalize: x / max(||x||, eps)
#pragma unroll
            for (int c = 0; c < 4; ++c) o[i][c] = acc[i][c] * inv;
        }
        pj_store_rows(H.raw, acc, row0, a.B, H.d_out, ty, tx);
        pj_store_rows(H.out, o, row0, a.B, H.d_out, ty, tx);
    }
}

__global__ void __launch_bounds__(PJ_NT) projector_bwd_kernel(const __grid_constant__ PjArgs a) {
    extern __shared__ __align__(16) float pj_sm[];
    const PjHead& H = a.h[blockIdx.y];
    float* w1s = pj_sm;                       // [d_hid][64]  as stored: dx = dh . W1
    float* w2s = w1s + PJ_D * PJ_D;           // [d_out][64]  as stored: dh = dp . W2
    float* xs = w2s + PJ_D * PJ_D;
    float* hs = xs + PJ_M * PJ_LD;
    float* dps = hs + PJ_M * PJ_LD;           // d loss / d raw
    float* dhs = dps + PJ_M * PJ_LD;          // d loss / d (W1 x + b1)
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    pj_load_weight(w1s, H.w1, H.d_hid, H.d_in, false, tid);
    pj_load_weight(w2s, H.w2, H.d_out, H.d_hid, false, tid);
    float gw1[4][4], gw2[4][4], gb1[4], gb2[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        gb1[i] = 0.f; gb2[i] = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) { gw1[i][c] = 0.f; gw2[i][c] = 0.f; }
    }
    const long long ntiles = (a.B + PJ_M - 1) / PJ_M;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const long long row0 = t * PJ_M;
        __syncthreads();
        pj_load_tile(xs, H.x, row0, a.B, H.d_in, tid);
        pj_load_tile(hs, H.hid, row0, a.B, H.d_hid, tid);
        // d raw from d out: out = raw / n, n = max(||raw||, eps); the norm carries gradient only where ||raw|| > eps
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const long long r = row0 + 4 * ty + i;
            float g[4], p[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const bool ok = r < a.B && 4 * tx + c < H.d_out;
                g[c] = ok ? __ldg(H.g_out + (size_t)r * H.d_out + 4 * tx + c) : 0.f;
                p[c] = ok ? __ldg(H.raw + (size_t)r * H.d_out + 4 * tx + c) : 0.f;
            }
            float ss = 0.f, dot = 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) { ss = fmaf(p[c], p[c], ss); dot = fmaf(g[c], p[c], dot); }
            ss = pj_sum16(ss); dot = pj_sum16(dot);
            const float nrm = sqrtf(ss);
            const float inv = 1.f / fmaxf(nrm, a.eps);
            const float k = nrm > a.eps ? dot * inv * inv * inv : 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) dps[(4 * ty + i) * PJ_LD + 4 * tx + c] = r < a.B ? g[c] * inv - p[c] * k : 0.f;
        }
        __syncthreads();
        {   // dh = (dp . W2) masked by the ReLU
            float acc[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[i][c] = 0.f;
            pj_gemm_rows(acc, dps, w2s, H.d_out, ty, tx);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int o = (4 * ty + i) * PJ_LD + 4 * tx + c;
                    dhs[o] = hs[o] > 0.f ? acc[i][c] : 0.f;
                }
        }
        __syncthreads();
        pj_gemm_cols(gw2, gb2, dps, hs, ty, tx);          // dW2[o][k] += dp[r][o] * h[r][k],  db2[o] += dp[r][o]
        pj_gemm_cols(gw1, gb1, dhs, xs, ty, tx);          // dW1[j][i] += dh[r][j] * x[r][i],  db1[j] += dh[r][j]
        if (H.dx) {
            float acc[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[i][c] = 0.f;
            pj_gemm_rows(acc, dhs, w1s, H.d_hid, ty, tx);
            pj_store_rows(H.dx, acc, row0, a.B, H.d_in, ty, tx);
        }
    }
    // this CTA's partial: [dw1 | db1 | dw2 | db2]
    float* P = H.scratch + (size_t)blockIdx.x * pj_partial_floats(H.d_in, H.d_hid, H.d_out);
    float* pw1 = P; float* pb1 = pw1 + H.d_hid * H.d_in; float* pw2 = pb1 + H.d_hid; float* pb2 = pw2 + H.d_out * H.d_hid;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int j = 4 * ty + i;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int k = 4 * tx + c;
            if (j < H.d_hid && k < H.d_in) pw1[j * H.d_in + k] = gw1[i][c];
            if (j < H.d_out && k < H.d_hid) pw2[j * H.d_hid + k] = gw2[i][c];
        }
        if (tx == 0) {
            if (j < H.d_hid) pb1[j] = gb1[i];
            if (j < H.d_out) pb2[j] = gb2[i];
        }
    }
}

// partials of `nctas` CTAs added in CTA order
__global__ void __launch_bounds__(PJ_NT) projector_reduce_kernel(const __grid_constant__ PjArgs a, int nctas) {
    const PjHead& H = a.h[blockIdx.y];
    const int n = pj_partial_floats(H.d_in, H.d_hid, H.d_out);
    const int e = blockIdx.x * PJ_NT + threadIdx.x;
    if (e >= n) return;
    float s = 0.f;
    for (int b = 0; b < nctas; ++b) s += H.scratch[(size_t)b * n + e];
    const int n1 = H.d_hid * H.d_in, n2 = n1 + H.d_hid, n3 = n2 + H.d_out * H.d_hid;
    if (e < n1) H.dw1[e] = s;
    else if (e < n2) H.db1[e - n1] = s;
    else if (e < n3) H.dw2[e - n2] = s;
    else H.db2[e - n3] = s;
}

static int pj_ctas(long long B) {
    const long long ntiles = (B + PJ_M - 1) / PJ_M;
    return (int)std::min<long long>(ntiles, sm_count());
}
static int pj_fill(PjArgs& a, const cfm_projector_t* heads, const cfm_projector_grads_t* grads, int64_t n_heads, int64_t B,
                   double eps) {
    CFM_REQUIRE(heads && n_heads >= 1 && n_heads <= PJ_MAX_HEADS, CFM_ERR_INVALID, "1..%d projection heads per call", PJ_MAX_HEADS);
    CFM_REQUIRE(B >= 1 && eps >= 0.0, CFM_ERR_INVALID, "bad batch size or eps");
    a.B = B; a.eps = (float)eps;
    for (int i = 0; i < n_heads; ++i) {
        const cfm_projector_t& p = heads[i];
        CFM_REQUIRE(p.d_in >= 1 && p.d_hid >= 1 && p.d_out >= 1, CFM_ERR_INVALID, "projector widths must be positive");
        CFM_REQUIRE(p.d_in <= PJ_D && p.d_hid <= PJ_D && p.d_out <= PJ_D, CFM_ERR_UNSUPPORTED,
                    "projector layers wider than %d are not built (%lld -> %lld -> %lld)", PJ_D, (long long)p.d_in,
                    (long long)p.d_hid, (long long)p.d_out);
        CFM_REQUIRE(p.w1 && p.b1 && p.w2 && p.b2 && p.x && p.hid && p.raw, CFM_ERR_INVALID, "null pointer in projector %d", i);
        PjHead& h = a.h[i];
        h = PjHead{};
        h.d_in = (int)p.d_in; h.d_hid = (int)p.d_hid; h.d_out = (int)p.d_out;
        h.w1 = p.w1; h.b1 = p.b1; h.w2 = p.w2; h.b2 = p.b2; h.x = p.x; h.hid = p.hid; h.raw = p.raw; h.out = p.out;
        if (grads) {
            const cfm_projector_grads_t& g = grads[i];
            CFM_REQUIRE(g.g_out && g.dw1 && g.db1 && g.dw2 && g.db2 && g.scratch, CFM_ERR_INVALID,
                        "null pointer in projector gradients %d", i);
            h.g_out = g.g_out; h.dx = g.dx; h.dw1 = g.dw1; h.db1 = g.db1; h.dw2 = g.dw2; h.db2 = g.db2; h.scratch = g.scratch;
        } else {
            CFM_REQUIRE(p.out, CFM_ERR_INVALID, "null output in projector %d", i);
        }
    }
    return CFM_OK;
}

}  // namespace cfm

extern "C" int64_t cfm_projector_scratch_floats(const cfm_projector_t* p, int64_t B) {
    if (!p || B < 1) return 0;
    return (int64_t)cfm::pj_ctas(B) * cfm::pj_partial_floats((int)p->d_in, (int)p->d_hid, (int)p->d_out);
}

extern "C" int cfm_projector_fwd(const cfm_projector_t* heads, int64_t n_heads, int64_t B, double eps, void* stream) {
    using namespace cfm;
    if (B == 0) return CFM_OK;
    PjArgs a{};
    if (int rc = pj_fill(a, heads, nullptr, n_heads, B, eps)) return rc;
    constexpr int smem = (2 * PJ_D * PJ_D + 2 * PJ_M * PJ_LD + 128) * 4;
    static bool attr = false;
    if (!attr) {
        CFM_CHECK_CUDA(cudaFuncSetAttribute(projector_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr = true;
    }
    projector_fwd_kernel<<<dim3(pj_ctas(B), (unsigned)n_heads), PJ_NT, smem, (cudaStream_t)stream>>>(a);
    CFM_LAUNCH_CHECK();
    return CFM_OK;
}

extern "C" int cfm_projector_bwd(const cfm_projector_t* heads, const cfm_projector_grads_t* grads, int64_t n_heads,
                                 int64_t B, double eps, void* stream) {
    using namespace cfm;
    CFM_REQUIRE(grads, CFM_ERR_INVALID, "null gradient descriptors");
    CFM_REQUIRE(B >= 1, CFM_ERR_INVALID, "the projector backward needs at least one row");
    PjArgs a{};
    if (int rc = pj_fill(a, heads, grads, n_heads, B, eps)) return rc;
    constexpr int smem = (2 * PJ_D * PJ_D + 4 * PJ_M * PJ_LD) * 4;
    static bool attr = false;
    if (!attr) {
        CFM_CHECK_CUDA(cudaFuncSetAttribute(projector_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr = true;
    }
    const int ctas = pj_ctas(B);
    projector_bwd_kernel<<<dim3(ctas, (unsigned)n_heads), PJ_NT, smem, (cudaStream_t)stream>>>(a);
    CFM_LAUNCH_CHECK();
    int nmax = 0;
    for (int i = 0; i < n_heads; ++i) nmax = std::max(nmax, pj_partial_floats(a.h[i].d_in, a.h[i].d_hid, a.h[i].d_out));
    projector_reduce_kernel<<<dim3((nmax + PJ_NT - 1) / PJ_NT, (unsigned)n_heads), PJ_NT, 0, (cudaStream_t)stream>>>(a, ctas);
    CFM_LAUNCH_CHECK();
    return CFM_OK;
}
