// Fused dense Adam step for the parameters of the two-tower model: ONE pass over
// (param, grad, exp_avg, exp_avg_sq) for every tensor of the model in a single launch, instead of the ~12
// multi-tensor passes of torch.optim.Adam(foreach, capturable).  At BASELINE config 4 the eleven 1M-row tables make
// the optimiser 1.4 GB of parameters: torch's passes move ~40 GB per step, this kernel 9.9 GB (7 floats per element).
//
// replaces torch.optim.Adam as used by training.py:32,55 (dense semantics: every row decays every step).  The
// arithmetic mirrors torch/optim/adam.py::_multi_tensor_adam, capturable branch, operation by operation and rounding
// by rounding (each foreach op rounds to fp32; contractions only where ATen's own functors allow them), with the step
// count kept on the device like torch's `state["step"]`.
#include "common.cuh"
#include <algorithm>

namespace cfm {

struct AdamTensor { float* p; const float* g; float* m; float* v; long long n; };
constexpr int ADAM_MAX_T = 64;               // tensors per launch: the records travel as kernel parameters (2.8 KB),
struct AdamBatch {                           // so a step needs no descriptor memory and is CUDA-graph capturable as is
    AdamTensor t[ADAM_MAX_T];
    int first_chunk[ADAM_MAX_T + 1];         // prefix sums of the per-tensor chunk counts
    int n;
};

constexpr int ADAM_CHUNK = 1 << 16;          // elements per (tensor, chunk) work item
constexpr int ADAM_NT = 256;

struct AdamScalars { float w1, beta2, w2, eps, inv_lr, beta1; };

// bias-correction scalars exactly as the capturable branch builds them (all fp32 tensor ops in torch)
__device__ __forceinline__ void adam_corrections(const AdamScalars& s, float step, float& step_size, float& bc2_sqrt) {
    float bc1 = __fsub_rn(powf(s.beta1, step), 1.f);          // _foreach_pow(beta1, steps); _foreach_sub_(.., 1)
    float bc2 = -__fsub_rn(powf(s.beta2, step), 1.f);         // ... ; _foreach_neg_
    // _foreach_div_(bc1, lr) multiplies by the scalar's reciprocal (ATen), then _foreach_reciprocal_: -lr / (1 - beta1^t)
    step_size = __frcp_rn(__fmul_rn(bc1, s.inv_lr));
    bc2_sqrt = __fsqrt_rn(bc2);                               // _foreach_sqrt_
}

template <int VARIANT>
__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, const AdamScalars& s, float step_size,
                                          float bc2_sqrt) {
    // exp_avg.lerp_(grad, 1 - beta1): ATen lerp, |weight| < 0.5 branch: self + weight * (end - self)
    const float diff = __fsub_rn(g, m);
    m = (VARIANT & 1) ? __fadd_rn(m, __fmul_rn(s.w1, diff)) : __fmaf_rn(s.w1, diff, m);
    // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value = 1 - beta2): self + value * t1 * t2
    const float v1 = __fmul_rn(v, s.beta2);
    const float gg = __fmul_rn(g, g);
    v = (VARIANT & 2) ? __fadd_rn(v1, __fmul_rn(s.w2, gg)) : __fmaf_rn(s.w2, gg, v1);
    // sqrt / bc2_sqrt + eps, / step_size, then param.addcdiv_(exp_avg, denom)
    float d = __fsqrt_rn(v);
    d = __fdiv_rn(d, bc2_sqrt);
    d = __fadd_rn(d, s.eps);
    d = __fdiv_rn(d, step_size);
    p = __fadd_rn(p, __fdiv_rn(m, d));
}

__global__ void adam_advance_step(float* step, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) step[i] = __fadd_rn(step[i], 1.f);
}

template <int VARIANT>
__global__ void __launch_bounds__(ADAM_NT) adam_kernel(const __grid_constant__ AdamBatch B, const float* __restrict__ step,
                                                       AdamScalars s) {
    float step_size, bc2_sqrt;
    adam_corrections(s, *step, step_size, bc2_sqrt);
    const int n_work = B.first_chunk[B.n];
    for (int wi = blockIdx.x; wi < n_work; wi += gridDim.x) {
        int lo_t = 0, hi_t = B.n - 1;                          // tensor owning chunk wi: last t with first_chunk[t] <= wi
        while (lo_t < hi_t) {
            const int mid = (lo_t + hi_t + 1) >> 1;
            if (B.first_chunk[mid] <= wi) lo_t = mid; else hi_t = mid - 1;
        }
        const AdamTensor& t = B.t[lo_t];
        const long long lo = (long long)(wi - B.first_chunk[lo_t]) * ADAM_CHUNK;
        const int cnt = (int)min((long long)ADAM_CHUNK, t.n - lo);
        float* p = t.p + lo; const float* g = t.g + lo; float* m = t.m + lo; float* v = t.v + lo;
        const bool al = ((((uintptr_t)p) | ((uintptr_t)g) | ((uintptr_t)m) | ((uintptr_t)v)) & 15) == 0;
        const int n4 = al ? cnt >> 2 : 0;
        // two float4 groups per thread and iteration: eight independent 16-byte loads in flight before the first use
        for (int i = threadIdx.x; i < n4; i += 2 * ADAM_NT) {
            const int j = i + ADAM_NT;
            const bool two = j < n4;
            float4 p4 = reinterpret_cast<float4*>(p)[i], m4 = reinterpret_cast<float4*>(m)[i],
                   v4 = reinterpret_cast<float4*>(v)[i];
            const float4 g4 = __ldcs(reinterpret_cast<const float4*>(g) + i);
            float4 q4 = p4, n4v = m4, w4 = v4, h4 = g4;
            if (two) {
                q4 = reinterpret_cast<float4*>(p)[j]; n4v = reinterpret_cast<float4*>(m)[j];
                w4 = reinterpret_cast<float4*>(v)[j]; h4 = __ldcs(reinterpret_cast<const float4*>(g) + j);
            }
            adam_elem<VARIANT>(p4.x, g4.x, m4.x, v4.x, s, step_size, bc2_sqrt);
            adam_elem<VARIANT>(p4.y, g4.y, m4.y, v4.y, s, step_size, bc2_sqrt);
            adam_elem<VARIANT>(p4.z, g4.z, m4.z, v4.z, s, step_size, bc2_sqrt);
            adam_elem<VARIANT>(p4.w, g4.w, m4.w, v4.w, s, step_size, bc2_sqrt);
            reinterpret_cast<float4*>(p)[i] = p4;
            reinterpret_cast<float4*>(m)[i] = m4;
            reinterpret_cast<float4*>(v)[i] = v4;
            if (two) {
                adam_elem<VARIANT>(q4.x, h4.x, n4v.x, w4.x, s, step_size, bc2_sqrt);
                adam_elem<VARIANT>(q4.y, h4.y, n4v.y, w4.y, s, step_size, bc2_sqrt);
                adam_elem<VARIANT>(q4.z, h4.z, n4v.z, w4.z, s, step_size, bc2_sqrt);
                adam_elem<VARIANT>(q4.w, h4.w, n4v.w, w4.w, s, step_size, bc2_sqrt);
                reinterpret_cast<float4*>(p)[j] = q4;
                reinterpret_cast<float4*>(m)[j] = n4v;
                reinterpret_cast<float4*>(v)[j] = w4;
            }
        }
        for (int i = 4 * n4 + threadIdx.x; i < cnt; i += ADAM_NT) {
            float pe = p[i], me = m[i], ve = v[i];
            adam_elem<VARIANT>(pe, g[i], me, ve, s, step_size, bc2_sqrt);
            p[i] = pe; m[i] = me; v[i] = ve;
        }
    }
}

}  // namespace cfm

using namespace cfm;

// tensors: HOST array of n_tensors records {param, grad, exp_avg, exp_avg_sq, numel}; step: device fp32 [n_steps]
// (torch keeps one state["step"] per parameter, all equal): every element is incremented first, element 0 is used.  `variant` selects the contraction pattern of the two fused ATen
// functors (0 = bit-equal to torch 2.11's kernels; 1-3 for the differential test).
extern "C" int cfm_adam_step(const cfm_adam_tensor_t* tensors, int64_t n_tensors, float* step, int64_t n_steps, double lr,
                             double beta1, double beta2, double eps, int64_t variant, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CFM_REQUIRE(tensors && step, CFM_ERR_INVALID, "null pointer");
    CFM_REQUIRE(n_tensors >= 1 && n_steps >= 1 && n_steps < (1 << 20), CFM_ERR_INVALID, "bad sizes");
    CFM_REQUIRE(lr > 0 && beta1 >= 0 && beta1 < 1 && beta2 >= 0 && beta2 < 1 && eps >= 0, CFM_ERR_INVALID,
                "bad Adam hyper-parameters");
    CFM_REQUIRE(1.0 - beta1 < 0.5, CFM_ERR_UNSUPPORTED, "beta1 <= 0.5 takes ATen's other lerp branch (not implemented)");
    AdamScalars s;
    s.w1 = (float)(1.0 - beta1); s.beta2 = (float)beta2; s.w2 = (float)(1.0 - beta2); s.eps = (float)eps;
    s.inv_lr = (float)(1.0 / lr); s.beta1 = (float)beta1;
    adam_advance_step<<<(int)((n_steps + 255) / 256), 256, 0, stream>>>(step, (int)n_steps);
    CFM_LAUNCH_CHECK();
    for (int64_t t0 = 0; t0 < n_tensors; t0 += ADAM_MAX_T) {
        AdamBatch B;
        B.n = (int)std::min<int64_t>(ADAM_MAX_T, n_tensors - t0);
        long long chunks = 0;
        for (int i = 0; i < B.n; ++i) {
            const cfm_adam_tensor_t& r = tensors[t0 + i];
            CFM_REQUIRE(r.param && r.grad && r.exp_avg && r.exp_avg_sq && r.numel >= 1, CFM_ERR_INVALID,
                        "bad tensor record %lld", (long long)(t0 + i));
            B.t[i] = AdamTensor{r.param, r.grad, r.exp_avg, r.exp_avg_sq, (long long)r.numel};
            B.first_chunk[i] = (int)chunks;
            chunks += (r.numel + ADAM_CHUNK - 1) / ADAM_CHUNK;
            CFM_REQUIRE(chunks < (1ll << 31), CFM_ERR_UNSUPPORTED, "too many elements for one launch");
        }
        for (int i = B.n; i <= ADAM_MAX_T; ++i) B.first_chunk[i] = (int)chunks;
        for (int i = B.n; i < ADAM_MAX_T; ++i) B.t[i] = AdamTensor{nullptr, nullptr, nullptr, nullptr, 0};
        const int grid = (int)std::min<long long>(chunks, (long long)sm_count() * 8);
        ProfScope prof(PROF_ADAM, stream);
        switch ((int)variant) {
            case 0: adam_kernel<0><<<grid, ADAM_NT, 0, stream>>>(B, step, s); break;
            case 1: adam_kernel<1><<<grid, ADAM_NT, 0, stream>>>(B, step, s); break;
            case 2: adam_kernel<2><<<grid, ADAM_NT, 0, stream>>>(B, step, s); break;
            default: adam_kernel<3><<<grid, ADAM_NT, 0, stream>>>(B, step, s); break;
        }
        CFM_LAUNCH_CHECK();
    }
    return CFM_OK;
}
