// Head kernels: cosine/MSE head of the two-tower model and the fused structural head.
// All are HBM-bound row-parallel kernels: coalesced loads, warp-shuffle row reductions, per-CTA partial
// sums finished by the last CTA in fixed order (deterministic, single launch).
#include "common.cuh"
#include <algorithm>

namespace cfm {

constexpr int HEAD_NT = 256;
constexpr int HEAD_MAX_CTAS = 1184;  // 8 CTAs per SM; partial buffers (>= 4096 floats) are sized for this many

// Sum `nv` per-thread values over the CTA (fixed tree), store as partial of this CTA; the last CTA to
// arrive sums all partials in CTA order into out[0..nv).  partial[0] is the arrival counter (must be 0 on
// entry; reset on exit), values start at partial[8].
template <int NV>
__device__ void cta_reduce_finish(float (&v)[NV], float* partial, float* const (&out)[NV], const float (&scale)[NV]) {
    __shared__ float sm_red[NV][HEAD_NT / 32];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        float s = warp_sum(v[i]);
        if (lane == 0) sm_red[i][warp] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            float s = 0.f;
            for (int w = 0; w < HEAD_NT / 32; ++w) s += sm_red[i][w];
            partial[8 + blockIdx.x * NV + i] = s;
        }
        __threadfence();
        unsigned prev = atomicAdd(reinterpret_cast<unsigned*>(partial), 1u);
        is_last = prev == gridDim.x - 1;
    }
    __syncthreads();
    if (is_last) {
        // every thread sums a strided subset of the per-CTA partials in a fixed order, then a fixed tree: the
        // result does not depend on which CTA happened to arrive last
        __threadfence();
        __shared__ float sm_fin[NV][HEAD_NT];
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            float s = 0.f;
            for (unsigned b = threadIdx.x; b < gridDim.x; b += HEAD_NT)
                s += reinterpret_cast<volatile float*>(partial)[8 + b * NV + i];
            sm_fin[i][threadIdx.x] = s;
        }
        __syncthreads();
        for (int off = HEAD_NT / 2; off > 0; off >>= 1) {
            if (threadIdx.x < off)
#pragma unroll
                for (int i = 0; i < NV; ++i) sm_fin[i][threadIdx.x] += sm_fin[i][threadIdx.x + off];
            __syncthreads();
        }
        if (threadIdx.x == 0) {
#pragma unroll
            for (int i = 0; i < NV; ++i)
                if (out[i]) *out[i] = sm_fin[i][0] * scale[i];
            *reinterpret_cast<unsigned*>(partial) = 0u;
        }
    }
}

// ------------------------------------------------------------------------------------------
// cosine head (model.py:79-87, contrastive.py:64,70,92-93, training.py:52)
// ------------------------------------------------------------------------------------------
struct CosArgs {
    const float *u, *v, *logit_scale;
    const float *target, *weights;        // fused weighted MSE (nullable)
    const float *d_score, *d_uhat, *d_vhat;   // backward inputs (nullable)
    const float *g_loss;                  // device scalar dL/d(loss) for the fused-MSE backward (nullable = 1)
    long long B; int D; float eps;
    float *score, *u_hat, *v_hat;         // forward outputs (nullable except score)
    float *du, *dv, *d_logit_scale, *loss, *partial;
    int mode;                             // 0 forward (+ optional MSE loss), 1 backward from d_score, 2 backward from MSE
};

// one warp per row; lanes stride over D
template <int MAXE>   // elements per lane kept in registers (D <= 32*MAXE)
__global__ void __launch_bounds__(HEAD_NT) cosine_head_kernel(const CosArgs a) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * (HEAD_NT / 32) + (threadIdx.x >> 5);
    const long long wstride = (long long)gridDim.x * (HEAD_NT / 32);
    const float scale = expf(a.logit_scale[0]);
    float acc[2] = {0.f, 0.f};   // [0] loss, [1] d_logit_scale
    const float gl = (a.mode == 2 && a.g_loss) ? a.g_loss[0] : 1.f;
    for (long long r = warp0; r < a.B; r += wstride) {
        float ue[MAXE], ve[MAXE];
        float nu2 = 0.f, nv2 = 0.f, dot = 0.f;
#pragma unroll
        for (int i = 0; i < MAXE; ++i) {
            int d = lane + 32 * i;
            ue[i] = d < a.D ? a.u[r * a.D + d] : 0.f;
            ve[i] = d < a.D ? a.v[r * a.D + d] : 0.f;
            nu2 = fmaf(ue[i], ue[i], nu2); nv2 = fmaf(ve[i], ve[i], nv2); dot = fmaf(ue[i], ve[i], dot);
        }
        nu2 = warp_sum(nu2); nv2 = warp_sum(nv2); dot = warp_sum(dot);
        const float nu = sqrtf(nu2), nv = sqrtf(nv2);
        const bool cl_u = a.eps > 0.f && nu < a.eps, cl_v = a.eps > 0.f && nv < a.eps;   // F.normalize clamp
        const float inu = 1.f / (cl_u ? a.eps : nu), inv = 1.f / (cl_v ? a.eps : nv);
        const float c = dot * inu * inv;
        const float s = scale * c;
        if (a.mode == 0) {
            if (lane == 0) {
                a.score[r] = s;
                if (a.target) { float diff = s - a.target[r]; acc[0] += a.weights[r] * diff * diff; }
            }
#pragma unroll
            for (int i = 0; i < MAXE; ++i) {
                int d = lane + 32 * i;
                if (d < a.D) {
                    if (a.u_hat) a.u_hat[r * a.D + d] = ue[i] * inu;
                    if (a.v_hat) a.v_hat[r * a.D + d] = ve[i] * inv;
                }
            }
        }
        if (a.mode == 0) continue;
        float ds;
        if (a.mode == 2) {
            ds = 2.f * a.weights[r] * (s - a.target[r]) * gl / (float)a.B;   // d mean(w (s-t)^2) / ds
            if (a.d_score) ds += a.d_score[r];
        } else {
            ds = a.d_score ? a.d_score[r] : 0.f;
        }
        if (lane == 0) acc[1] += ds * s;
        // g wrt u_hat = ds*scale*v_hat + d_uhat ; du = (g - u_hat (u_hat . g)) / |u|   (or g/eps when clamped)
        float gu[MAXE], gv[MAXE], pu = 0.f, pv = 0.f;
        const float k = ds * scale;
#pragma unroll
        for (int i = 0; i < MAXE; ++i) {
            int d = lane + 32 * i;
            float uh = ue[i] * inu, vh = ve[i] * inv;
            gu[i] = k * vh + ((a.d_uhat && d < a.D) ? a.d_uhat[r * a.D + d] : 0.f);
            gv[i] = k * uh + ((a.d_vhat && d < a.D) ? a.d_vhat[r * a.D + d] : 0.f);
            pu = fmaf(uh, gu[i], pu); pv = fmaf(vh, gv[i], pv);
        }
        pu = warp_sum(pu); pv = warp_sum(pv);
#pragma unroll
        for (int i = 0; i < MAXE; ++i) {
            int d = lane + 32 * i;
            if (d < a.D) {
                float uh = ue[i] * inu, vh = ve[i] * inv;
                a.du[r * a.D + d] = cl_u ? gu[i] * inu : (gu[i] - uh * pu) * inu;
                a.dv[r * a.D + d] = cl_v ? gv[i] * inv : (gv[i] - vh * pv) * inv;
            }
        }
    }
    if (a.mode != 0 || a.loss) {
        float* const outs[2] = {a.loss, a.d_logit_scale};
        const float sc[2] = {1.f / (float)a.B, 1.f};
        cta_reduce_finish<2>(acc, a.partial, outs, sc);
    }
}

// D % 4 == 0, D <= 128: LPR lanes per row (16 for D <= 64, else 32), one float4 per lane, so a warp covers 32/LPR rows
// with 16-byte accesses and log2(LPR) shuffle steps per reduction.  Same arithmetic as cosine_head_kernel.
__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
    return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}
template <int LPR>
__global__ void __launch_bounds__(HEAD_NT) cosine_head_kernel_v4(const CosArgs a) {
    constexpr int RPW = 32 / LPR;
    const int lane = threadIdx.x & 31, sub = lane / LPR, q = lane % LPR;
    const long long wbase = ((long long)blockIdx.x * (HEAD_NT / 32) + (threadIdx.x >> 5)) * RPW;
    const long long wstride = (long long)gridDim.x * (HEAD_NT / 32) * RPW;
    const int D4 = a.D >> 2;
    const bool on = q < D4;
    const float scale = expf(a.logit_scale[0]);
    float acc[2] = {0.f, 0.f};   // [0] loss, [1] d_logit_scale
    const float gl = (a.mode == 2 && a.g_loss) ? a.g_loss[0] : 1.f;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (long long rb = wbase; rb < a.B; rb += wstride) {
        const long long r = rb + sub;
        const bool live = r < a.B, ld = live && on;
        const float4 ue = ld ? *(reinterpret_cast<const float4*>(a.u + r * a.D) + q) : z4;
        const float4 ve = ld ? *(reinterpret_cast<const float4*>(a.v + r * a.D) + q) : z4;
        float nu2 = dot4(ue, ue), nv2 = dot4(ve, ve), dot = dot4(ue, ve);
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) {
            nu2 += __shfl_xor_sync(FULL, nu2, o);
            nv2 += __shfl_xor_sync(FULL, nv2, o);
            dot += __shfl_xor_sync(FULL, dot, o);
        }
        const float nu = sqrtf(nu2), nv = sqrtf(nv2);
        const bool cl_u = a.eps > 0.f && nu < a.eps, cl_v = a.eps > 0.f && nv < a.eps;   // F.normalize clamp
        const float inu = 1.f / (cl_u ? a.eps : nu), inv = 1.f / (cl_v ? a.eps : nv);
        const float s = scale * (dot * inu * inv);
        const float4 uh = make_float4(ue.x * inu, ue.y * inu, ue.z * inu, ue.w * inu);
        const float4 vh = make_float4(ve.x * inv, ve.y * inv, ve.z * inv, ve.w * inv);
        if (a.mode == 0) {
            if (live && q == 0) {
                a.score[r] = s;
                if (a.target) { float diff = s - a.target[r]; acc[0] += a.weights[r] * diff * diff; }
            }
            if (ld) {
                if (a.u_hat) *(reinterpret_cast<float4*>(a.u_hat + r * a.D) + q) = uh;
                if (a.v_hat) *(reinterpret_cast<float4*>(a.v_hat + r * a.D) + q) = vh;
            }
            continue;
        }
        float ds = 0.f;
        if (live) {
            if (a.mode == 2) {
                ds = 2.f * a.weights[r] * (s - a.target[r]) * gl / (float)a.B;   // d mean(w (s-t)^2) / ds
                if (a.d_score) ds += a.d_score[r];
            } else {
                ds = a.d_score ? a.d_score[r] : 0.f;
            }
        }
        if (live && q == 0) acc[1] += ds * s;
        // g wrt u_hat = ds*scale*v_hat + d_uhat ; du = (g - u_hat (u_hat . g)) / |u|   (or g/eps when clamped)
        const float k = ds * scale;
        const float4 eu = (a.d_uhat && ld) ? *(reinterpret_cast<const float4*>(a.d_uhat + r * a.D) + q) : z4;
        const float4 ev = (a.d_vhat && ld) ? *(reinterpret_cast<const float4*>(a.d_vhat + r * a.D) + q) : z4;
        const float4 gu = make_float4(fmaf(k, vh.x, eu.x), fmaf(k, vh.y, eu.y), fmaf(k, vh.z, eu.z), fmaf(k, vh.w, eu.w));
        const float4 gv = make_float4(fmaf(k, uh.x, ev.x), fmaf(k, uh.y, ev.y), fmaf(k, uh.z, ev.z), fmaf(k, uh.w, ev.w));
        float pu = dot4(uh, gu), pv = dot4(vh, gv);
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) {
            pu += __shfl_xor_sync(FULL, pu, o);
            pv += __shfl_xor_sync(FULL, pv, o);
        }
        if (ld) {
            const float4 du = cl_u ? make_float4(gu.x * inu, gu.y * inu, gu.z * inu, gu.w * inu)
                                   : make_float4((gu.x - uh.x * pu) * inu, (gu.y - uh.y * pu) * inu,
                                                 (gu.z - uh.z * pu) * inu, (gu.w - uh.w * pu) * inu);
            const float4 dv = cl_v ? make_float4(gv.x * inv, gv.y * inv, gv.z * inv, gv.w * inv)
                                   : make_float4((gv.x - vh.x * pv) * inv, (gv.y - vh.y * pv) * inv,
                                                 (gv.z - vh.z * pv) * inv, (gv.w - vh.w * pv) * inv);
            *(reinterpret_cast<float4*>(a.du + r * a.D) + q) = du;
            *(reinterpret_cast<float4*>(a.dv + r * a.D) + q) = dv;
        }
    }
    if (a.mode != 0 || a.loss) {
        float* const outs[2] = {a.loss, a.d_logit_scale};
        const float sc[2] = {1.f / (float)a.B, 1.f};
        cta_reduce_finish<2>(acc, a.partial, outs, sc);
    }
}

static int launch_cosine(const CosArgs& a, cudaStream_t stream) {
    CFM_REQUIRE(a.B >= 1 && a.D >= 1 && a.D <= 256, CFM_ERR_UNSUPPORTED, "cosine head supports 1 <= D <= 256 (got %d)", a.D);
    int ctas = (int)std::min<long long>((a.B + HEAD_NT / 32 - 1) / (HEAD_NT / 32), HEAD_MAX_CTAS);
    ProfScope prof(PROF_HEAD, stream);
    const bool al16 = (((uintptr_t)a.u | (uintptr_t)a.v | (uintptr_t)a.du | (uintptr_t)a.dv | (uintptr_t)a.u_hat |
                        (uintptr_t)a.v_hat | (uintptr_t)a.d_uhat | (uintptr_t)a.d_vhat) & 15) == 0;
    if ((a.D & 3) == 0 && a.D <= 128 && al16) {
        const int rpw = a.D <= 64 ? 2 : 1, wpc = HEAD_NT / 32;
        ctas = (int)std::min<long long>((a.B + (long long)wpc * rpw - 1) / ((long long)wpc * rpw), HEAD_MAX_CTAS);
        if (a.D <= 64) cosine_head_kernel_v4<16><<<ctas, HEAD_NT, 0, stream>>>(a);
        else cosine_head_kernel_v4<32><<<ctas, HEAD_NT, 0, stream>>>(a);
    } else if (a.D <= 64) cosine_head_kernel<2><<<ctas, HEAD_NT, 0, stream>>>(a);
    else if (a.D <= 128) cosine_head_kernel<4><<<ctas, HEAD_NT, 0, stream>>>(a);
    else cosine_head_kernel<8><<<ctas, HEAD_NT, 0, stream>>>(a);
    CFM_LAUNCH_CHECK();
    return CFM_OK;
}

// ------------------------------------------------------------------------------------------
// structural head (structural_model.py:130-141, structural_training.py:75-77)
// ------------------------------------------------------------------------------------------
struct StructArgs {
    const float *c_logits, *f_logits, *A, *t_ceo, *t_firm, *d_match;
    long long B; float kl_scale;
    float *match, *loss, *d_c, *d_f, *partial;
};

__device__ __forceinline__ void softmax5(const float* z, float* p, float& lse) {
    float m = fmaxf(fmaxf(fmaxf(z[0], z[1]), fmaxf(z[2], z[3])), z[4]);
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 5; ++k) { p[k] = expf(z[k] - m); s += p[k]; }
    float inv = 1.f / s;
#pragma unroll
    for (int k = 0; k < 5; ++k) p[k] *= inv;
    lse = m + logf(s);
}

__global__ void __launch_bounds__(HEAD_NT) structural_head_kernel(const StructArgs a) {
    // tiles of HEAD_NT rows: [rows,5] blocks staged through shared memory for coalesced traffic
    __shared__ __align__(16) float sc[HEAD_NT * 5], sf[HEAD_NT * 5], st_c[HEAD_NT * 5], st_f[HEAD_NT * 5];
    __shared__ float sA[25];
    if (threadIdx.x < 25) sA[threadIdx.x] = a.A[threadIdx.x];
    float acc[1] = {0.f};
    const bool has_t = a.t_ceo != nullptr, has_grad = a.d_c != nullptr;
    const long long ntiles = (a.B + HEAD_NT - 1) / HEAD_NT;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long row0 = tile * HEAD_NT;
        const int rows = (int)min((long long)HEAD_NT, a.B - row0);
        __syncthreads();
        for (int i = threadIdx.x; i < rows * 5; i += HEAD_NT) {
            sc[i] = a.c_logits[row0 * 5 + i];
            sf[i] = a.f_logits[row0 * 5 + i];
            if (has_t) { st_c[i] = a.t_ceo[row0 * 5 + i]; st_f[i] = a.t_firm[row0 * 5 + i]; }
        }
        __syncthreads();
        const int r = threadIdx.x;
        float dc[5], df[5];
        if (r < rows) {
            float zc[5], zf[5], pi[5], q[5], lse_c, lse_f;
#pragma unroll
            for (int k = 0; k < 5; ++k) { zc[k] = sc[r * 5 + k]; zf[k] = sf[r * 5 + k]; }
            softmax5(zc, pi, lse_c);
            softmax5(zf, q, lse_f);
            float Aq[5], Atp[5], m = 0.f;   // (A q)_a, (A^T pi)_b
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                float s = 0.f, t = 0.f;
#pragma unroll
                for (int j = 0; j < 5; ++j) { s = fmaf(sA[i * 5 + j], q[j], s); t = fmaf(sA[j * 5 + i], pi[j], t); }
                Aq[i] = s; Atp[i] = t;
            }
#pragma unroll
            for (int i = 0; i < 5; ++i) m = fmaf(pi[i], Aq[i], m);
            a.match[row0 + r] = m;
            const float dm = a.d_match ? a.d_match[row0 + r] : 0.f;
#pragma unroll
            for (int k = 0; k < 5; ++k) { dc[k] = dm * pi[k] * (Aq[k] - m); df[k] = dm * q[k] * (Atp[k] - m); }
            if (has_t) {
                float kl = 0.f, stc = 0.f, stf = 0.f;
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    float tc = st_c[r * 5 + k], tf = st_f[r * 5 + k];
                    // KLDivLoss pointwise: xlogy(t, t) - t * log_softmax  (0 where t == 0)
                    if (tc > 0.f) kl += tc * (logf(tc) - (zc[k] - lse_c));
                    if (tf > 0.f) kl += tf * (logf(tf) - (zf[k] - lse_f));
                    stc += tc; stf += tf;
                }
                acc[0] += kl;
                const float g = a.kl_scale / (float)a.B;
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    dc[k] += g * (pi[k] * stc - st_c[r * 5 + k]);
                    df[k] += g * (q[k] * stf - st_f[r * 5 + k]);
                }
            }
        }
        if (has_grad) {
            __syncthreads();
            if (r < rows) {
#pragma unroll
                for (int k = 0; k < 5; ++k) { sc[r * 5 + k] = dc[k]; sf[r * 5 + k] = df[k]; }
            }
            __syncthreads();
            for (int i = threadIdx.x; i < rows * 5; i += HEAD_NT) {
                a.d_c[row0 * 5 + i] = sc[i];
                a.d_f[row0 * 5 + i] = sf[i];
            }
        }
    }
    float* const outs[1] = {a.loss};
    const float scl[1] = {1.f / (float)a.B};
    cta_reduce_finish<1>(acc, a.partial, outs, scl);
}

}  // namespace cfm

using namespace cfm;

extern "C" int cfm_cosine_head_fwd(const float* u, const float* v, const float* logit_scale, int64_t B, int64_t D,
                                   double eps, float* score, float* u_hat, float* v_hat, const float* target,
                                   const float* weights, float* loss, float* partial, void* stream) {
    CFM_REQUIRE(u && v && logit_scale && score, CFM_ERR_INVALID, "null pointer");
    CFM_REQUIRE(!loss || (target && weights && partial), CFM_ERR_INVALID, "fused loss needs target, weights, partial");
    if (B == 0) return CFM_OK;
    CosArgs a{};
    a.u = u; a.v = v; a.logit_scale = logit_scale; a.B = B; a.D = (int)D; a.eps = (float)eps;
    a.score = score; a.u_hat = u_hat; a.v_hat = v_hat; a.mode = 0;
    if (loss) { a.target = target; a.weights = weights; a.loss = loss; a.partial = partial; }
    return launch_cosine(a, (cudaStream_t)stream);
}

extern "C" int cfm_cosine_head_bwd(const float* u, const float* v, const float* logit_scale, const float* d_score,
                                   const float* d_uhat, const float* d_vhat, const float* target,
                                   const float* weights, const float* g_loss, int64_t B, int64_t D, double eps,
                                   float* du, float* dv, float* d_logit_scale, float* partial, void* stream) {
    CFM_REQUIRE(u && v && logit_scale && du && dv && d_logit_scale && partial, CFM_ERR_INVALID, "null pointer");
    CFM_REQUIRE((target == nullptr) == (weights == nullptr), CFM_ERR_INVALID, "give target and weights together");
    if (B == 0) return CFM_OK;
    CosArgs a{};
    a.u = u; a.v = v; a.logit_scale = logit_scale; a.d_score = d_score; a.d_uhat = d_uhat; a.d_vhat = d_vhat;
    a.target = target; a.weights = weights; a.g_loss = g_loss;
    a.B = B; a.D = (int)D; a.eps = (float)eps; a.du = du; a.dv = dv; a.d_logit_scale = d_logit_scale;
    a.partial = partial; a.mode = target ? 2 : 1;
    return launch_cosine(a, (cudaStream_t)stream);
}

extern "C" int cfm_structural_head(const float* c_logits, const float* f_logits, const float* A,
                                   const float* target_ceo, const float* target_firm, const float* d_match, int64_t B,
                                   double kl_scale, float* match, float* loss, float* d_c_logits, float* d_f_logits,
                                   float* partial, void* stream) {
    CFM_REQUIRE(c_logits && f_logits && A && match && partial, CFM_ERR_INVALID, "null pointer");
    CFM_REQUIRE((target_ceo == nullptr) == (target_firm == nullptr), CFM_ERR_INVALID, "give both targets or neither");
    CFM_REQUIRE((d_c_logits == nullptr) == (d_f_logits == nullptr), CFM_ERR_INVALID, "give both logit grads or neither");
    if (B == 0) return CFM_OK;
    StructArgs a{};
    a.c_logits = c_logits; a.f_logits = f_logits; a.A = A; a.t_ceo = target_ceo; a.t_firm = target_firm;
    a.d_match = d_match; a.B = B; a.kl_scale = (float)kl_scale; a.match = match; a.loss = loss;
    a.d_c = d_c_logits; a.d_f = d_f_logits; a.partial = partial;
    int ctas = (int)std::min<long long>((B + HEAD_NT - 1) / HEAD_NT, HEAD_MAX_CTAS);
    ProfScope prof(PROF_HEAD, (cudaStream_t)stream);
    structural_head_kernel<<<ctas, HEAD_NT, 0, (cudaStream_t)stream>>>(a);
    CFM_LAUNCH_CHECK();
    return CFM_OK;
}
