// Tower kernels: categorical-embedding gather + numeric concat + 3-layer MLP, forward and backward, one Linear per
// launch ("stage"), both towers in one grid (blockIdx.y).
//
// replaces ceo_firm_matching/model.py:69-76 and structural_model.py:120-127 (+ their autograd).
//
// Layout of one stage launch: persistent CTAs (two resident per SM) walk row tiles (64 rows; 32 in the stage-1
// backward so that the weight matrix plus the gathered tile fit twice per SM).  The stage's weight matrix lives in
// shared memory for the CTA's lifetime; each tile's input activations are (re)built in shared memory (stage 1:
// embedding rows gathered with cp.async + numerics; stage >1: BN/ReLU/dropout of the previous stage's raw output)
// and multiplied on the tensor cores: mma.sync.m16n8k8 TF32 with ldmatrix fragment loads, either one pass
// (precision 1) or three error-compensated passes a_lo*b_hi + a_hi*b_lo + a_hi*b_hi (precision 0, fp32-class).
// Train-mode BatchNorm needs whole-batch statistics between stages, hence the stage-per-launch split: every CTA
// writes (count, mean, M2) partials (two-pass per tile, Chan-merged), a finalize kernel merges them in a fixed order
// -> bitwise reproducible.  Backward mirrors this: per-CTA weight-gradient accumulators stay in registers over all
// tiles and are written once, then reduced in fixed CTA order.
#include "common.cuh"
#include "tower_types.cuh"
#include "tower_tc.cuh"
#include <algorithm>

namespace cfm {

// ------------------------------------------------------------------------------------------
// shared-memory carve-up (same arithmetic on host and device)
// ------------------------------------------------------------------------------------------
struct FwdSmem {
    int lda, ws, as, bn, red, tmean, rmean, rm2, idx, total_floats;
};
__host__ __device__ inline FwdSmem fwd_smem(int K, int N, int n_tab, const MmaPlan& gp, int tm, bool stage1) {
    FwdSmem s;
    s.lda = pad_ld(ceil8(K + 1));
    int o = 0;
    s.ws = o; o += gp.wrows * s.lda;
    s.as = o; o += tm * s.lda;
    s.bn = o; o += stage1 ? 0 : 4 * ((K + 3) & ~3);
    // single-pass layers: the statistics scratch aliases the input tile (dead once every warp has left the MMAs);
    // with several column passes the tile is still needed, so the scratch gets its own space
    if (gp.passes == 1) s.red = s.as;
    else { s.red = o; o += (tm / 4) * 128; }
    s.tmean = o; o += 128;
    s.rmean = o; o += gp.wrows + 128;
    s.rm2 = o; o += gp.wrows + 128;
    s.idx = o; o += tm * (n_tab > 0 ? n_tab : 1);
    s.total_floats = (o + 3) & ~3;
    return s;
}

struct BwdSmem {
    int lda, ldg, ldgt, ldxh, wt, as, gs, gt, xh, bna, bng, red, rs, idx, total_floats;
};
__host__ __device__ inline BwdSmem bwd_smem(int K, int N, int n_tab, int a_bn, int need_dx, const MmaPlan& gx,
                                            int tm, bool stage1) {
    BwdSmem s;
    s.lda = pad_ld_t(K + 1);            // A tile is the "transposed" B operand of the dW product
    s.ldg = pad_ld(ceil8(N));
    s.ldgt = pad_ld(tm);
    s.ldxh = pad_ld(K);
    int o = 0;
    s.wt = o; o += need_dx ? gx.wrows * s.ldg : 0;
    s.as = o; o += tm * s.lda;
    s.gs = o; o += tm * s.ldg;
    s.gt = o; o += ((N + 15) & ~15) * s.ldgt;
    s.xh = o; o += a_bn ? tm * s.ldxh : 0;
    s.bna = o; o += stage1 ? 0 : 4 * ((K + 3) & ~3);
    s.bng = o; o += 5 * ((N + 3) & ~3);
    s.red = o; o += a_bn ? 2 * (tm / 4) * 128 : 0;
    s.rs = o; o += a_bn ? 2 * (gx.wrows + 128) : 0;
    s.idx = o; o += tm * (n_tab > 0 ? n_tab : 1);
    s.total_floats = (o + 3) & ~3;
    return s;
}

// ------------------------------------------------------------------------------------------
// input tile builders
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void stage_bn_params(const ActSrc& a, int K, float* sm_bn) {
    int Kp = (K + 3) & ~3;
    if (a.bn_mode == 0) return;
    for (int c = threadIdx.x; c < K; c += NT) {
        float m = a.mean[c], s = a.var_or_istd[c];
        sm_bn[c] = m;
        sm_bn[Kp + c] = a.bn_mode == 2 ? rsqrtf(s + BN_EPS) : s;
        sm_bn[2 * Kp + c] = a.gamma[c];
        sm_bn[3 * Kp + c] = a.beta[c];
    }
}

// Build the [TM, K] input tile (+ ones column at K, zero pad to lda); optionally the x-hat tile.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
                 : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
                 : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Build the [TM, K] input tile (+ ones column at K, zero pad to lda); optionally the x-hat tile.
// Memory-level parallelism is what matters here (every row of a 1M-row table is an HBM miss): the gather is issued
// as asynchronous 16-byte global->shared copies, all in flight at once, and the activation rebuild loads four
// float4 per thread before touching any of them.  The caller's __syncthreads() publishes the tile.
__device__ void build_input_tile(const InputDesc& in, const DropCtx& drop, long long row0, int rows_valid, int tm,
                                 float* As, int lda, float* Xh, int ldxh, const float* sm_bn, int* sm_idx, int* err) {
    const int K = in.K;
    const int tid = threadIdx.x;
    if (in.stage == 1) {
        const GatherSrc& g = in.g;
        const int KE = g.n_tab * g.E;
        for (int i = tid; i < tm * g.n_tab; i += NT) {
            int r = i / g.n_tab, t = i - r * g.n_tab;
            long long idx = 0;
            if (r < rows_valid) {
                idx = g.x_cat[(row0 + r) * g.n_tab + t];
                if (idx < 0 || idx >= g.tab_rows[t]) {
                    if (err) atomicOr(err, CFM_FLAG_INDEX_OOB);
                    idx = 0;
                }
            }
            sm_idx[i] = (int)idx;
        }
        for (int i = tid; i < tm * g.n_num; i += NT) {              // numerics do not depend on the indices
            int r = i / g.n_num, j = i - r * g.n_num;
            if (r < rows_valid) cp_async4(As + r * lda + KE + j, g.x_num + (row0 + r) * g.n_num + j);
            else As[r * lda + KE + j] = 0.f;
        }
        __syncthreads();
        if ((g.E & 3) == 0) {
            const int E4 = g.E >> 2;
            const int items = tm * g.n_tab * E4;
            for (int i = tid; i < items; i += NT) {
                int q = i % E4;
                int rt = i / E4;
                int t = rt % g.n_tab, r = rt / g.n_tab;
                float* dst = As + r * lda + t * g.E + 4 * q;
                if (r < rows_valid) cp_async16(dst, g.tab[t] + (size_t)sm_idx[rt] * g.E + 4 * q);
                else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        } else {
            const int items = tm * KE;
            for (int i = tid; i < items; i += NT) {
                int e = i % g.E;
                int rt = i / g.E;
                int t = rt % g.n_tab, r = rt / g.n_tab;
                float* dst = As + r * lda + t * g.E + e;
                if (r < rows_valid) cp_async4(dst, g.tab[t] + (size_t)sm_idx[rt] * g.E + e);
                else *dst = 0.f;
            }
        }
    } else {
        const ActSrc& a = in.a;
        const int K4 = (K + 3) >> 2, Kp = K4 << 2;
        const bool vec = (K & 3) == 0;
        constexpr int U = 4;
        for (int base = tid; base < tm * K4; base += NT * U) {
            float v[U][4];
            // phase 1: all loads of this batch
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = base + u * NT;
                const int r = i / K4, c4 = i - r * K4;
                v[u][0] = v[u][1] = v[u][2] = v[u][3] = 0.f;
                if (i < tm * K4 && r < rows_valid) {
                    const float* hp = a.h + (size_t)(row0 + r) * K + 4 * c4;
                    if (vec) {
                        float4 t4 = *reinterpret_cast<const float4*>(hp);
                        v[u][0] = t4.x; v[u][1] = t4.y; v[u][2] = t4.z; v[u][3] = t4.w;
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (4 * c4 + e < K) v[u][e] = hp[e];
                    }
                }
            }
            // phase 2: BN / ReLU / dropout and the smem stores
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = base + u * NT;
                if (i >= tm * K4) continue;
                const int r = i / K4, c4 = i - r * K4;
                const bool valid = r < rows_valid;
                float xh[4] = {0.f, 0.f, 0.f, 0.f};
                if (valid) {
                    Philox4 w = {0, 0, 0, 0};
                    if (drop.active) w = drop_words(drop, row0 + r, c4);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        int c = 4 * c4 + e;
                        if (c < K) {
                            float t = v[u][e];
                            if (a.bn_mode) {
                                xh[e] = (t - sm_bn[c]) * sm_bn[Kp + c];
                                t = xh[e] * sm_bn[2 * Kp + c] + sm_bn[3 * Kp + c];
                            }
                            t = fmaxf(t, 0.f);
                            if (drop.active) t = drop_keep(drop, w, e) ? t * drop.inv_keep : 0.f;
                            v[u][e] = t;
                        } else {
                            v[u][e] = 0.f;
                        }
                    }
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    int c = 4 * c4 + e;
                    if (c < K) {
                        As[r * lda + c] = v[u][e];
                        if (Xh) Xh[r * ldxh + c] = xh[e];
                    }
                }
            }
        }
    }
    // ones column (folds the bias / bias gradient into the GEMMs) and zero padding
    for (int i = tid; i < tm * (lda - K); i += NT) {
        int r = i / (lda - K), c = K + (i - r * (lda - K));
        As[r * lda + c] = (c == K && r < rows_valid) ? 1.f : 0.f;
    }
    if (in.stage == 1) cp_async_wait_all();
}

// ------------------------------------------------------------------------------------------
// tensor-core GEMMs on shared-memory operands (mma.sync m16n8k8, 3xTF32 error-compensated)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// four 8x8 b16 matrices == four 8-row x 4-float blocks: one instruction fetches a whole TF32 A fragment
// (or the B fragments of two n-tiles); lane l supplies the 16-byte row address of matrix l/8, row l%8
__device__ __forceinline__ void ldmatrix_x4(const float* p, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void ldmatrix_x2(const float* p, uint32_t (&r)[2]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];"
                 : "=r"(r[0]), "=r"(r[1])
                 : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
template <bool EXACT>
__device__ __forceinline__ void split_frag(uint32_t raw, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(__uint_as_float(raw)));
    if (EXACT) lo = __float_as_uint(__uint_as_float(raw) - __uint_as_float(hi));
}
template <bool EXACT>
__device__ __forceinline__ void mma_acc(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                        const uint32_t (&bh)[2], const uint32_t (&bl)[2]) {
    if (EXACT) { mma_tf32(c, al, bh); mma_tf32(c, ah, bl); }
    mma_tf32(c, ah, bh);
}

// acc[mi][ni][.] = X[r0 + mi*16 .. +15][:] . W[n0 + ni*8 .. +7][:]^T over k8n steps of 8 columns.
// EXACT: 3xTF32 error-compensated (fp32-class result); otherwise one TF32 pass.  Leading dimensions are 4 (mod 8)
// floats, so the eight 16-byte rows of every ldmatrix phase fall into distinct bank groups.
template <int NI, bool EXACT>
__device__ __forceinline__ void mma_rows(const float* __restrict__ Xs, int ldx, const float* __restrict__ Ws, int ldw,
                                         int k8n, int r0, int n0, int lane, float (&acc)[2][NI][4]) {
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < NI; ++ni)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[mi][ni][c] = 0.f;
    const int mat = lane >> 3, rin = lane & 7;
    // A: matrices (rows 0-7 | 8-15) x (cols 0-3 | 4-7);  B (two n-tiles per x4): (n-tile 0 | 1) x (cols 0-3 | 4-7)
    const float* xp = Xs + (r0 + rin + (mat & 1) * 8) * ldx + (mat >> 1) * 4;
    const float* wp = Ws + (n0 + rin + (NI >= 2 ? (mat >> 1) * 8 : 0)) * ldw + (mat & 1) * 4;
#pragma unroll 2
    for (int k8 = 0; k8 < k8n; ++k8) {
        uint32_t ah[2][4], al[2][4], bh[NI][2], bl[NI][2];
#pragma unroll
        for (int mi = 0; mi < 2; ++mi) {
            uint32_t raw[4];
            ldmatrix_x4(xp + mi * 16 * ldx + k8 * 8, raw);
#pragma unroll
            for (int e = 0; e < 4; ++e) split_frag<EXACT>(raw[e], ah[mi][e], al[mi][e]);
        }
        if (NI == 1) {
            uint32_t raw[2];
            ldmatrix_x2(wp + k8 * 8, raw);
            split_frag<EXACT>(raw[0], bh[0][0], bl[0][0]);
            split_frag<EXACT>(raw[1], bh[0][1], bl[0][1]);
        } else {
#pragma unroll
            for (int np = 0; np < NI / 2; ++np) {
                uint32_t raw[4];
                ldmatrix_x4(wp + np * 16 * ldw + k8 * 8, raw);
                split_frag<EXACT>(raw[0], bh[2 * np][0], bl[2 * np][0]);
                split_frag<EXACT>(raw[1], bh[2 * np][1], bl[2 * np][1]);
                split_frag<EXACT>(raw[2], bh[2 * np + 1][0], bl[2 * np + 1][0]);
                split_frag<EXACT>(raw[3], bh[2 * np + 1][1], bl[2 * np + 1][1]);
            }
        }
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int ni = 0; ni < NI; ++ni) mma_acc<EXACT>(acc[mi][ni], ah[mi], al[mi], bh[ni], bl[ni]);
    }
}

struct FwdCtx {
    const float *As, *Ws;
    float *red, *tmean, *rmean, *rm2, *hout;
    int lda, k8n, N;
    bool stats;
    int wcn;              // column groups of warps (4 for 64-row tiles, 8 for 32-row tiles); row groups = 8 / wcn
};

// one column pass (<= 128 columns) of a forward tile: GEMM, store raw output, optional (count, mean, M2) statistics
template <int NI, bool EXACT, int WCN>
__device__ __forceinline__ void fwd_pass(const FwdCtx& C, int pass, int cols, long long row0, int rows_valid, float run_cnt) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wr = warp / WCN, wc = warp % WCN, g = lane >> 2, t = lane & 3;
    constexpr int nred = 8 * (8 / WCN);                   // partial rows per column in `red`
    const int n0 = pass * 128 + wc * 8 * NI;
    const bool active = wc * 8 * NI < cols;
    float acc[2][NI][4];
    if (active) {
        mma_rows<NI, EXACT>(C.As, C.lda, C.Ws, C.lda, C.k8n, wr * 32, n0, lane, acc);
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int r = wr * 32 + mi * 16 + g + 8 * h;
                if (r < rows_valid) {
#pragma unroll
                    for (int ni = 0; ni < NI; ++ni) {
                        const int c = n0 + ni * 8 + 2 * t;
                        float* dst = C.hout + (size_t)(row0 + r) * C.N + c;
                        if (c + 1 < C.N && !(C.N & 1)) {
                            *reinterpret_cast<float2*>(dst) = make_float2(acc[mi][ni][2 * h], acc[mi][ni][2 * h + 1]);
                        } else {
                            if (c < C.N) dst[0] = acc[mi][ni][2 * h];
                            if (c + 1 < C.N) dst[1] = acc[mi][ni][2 * h + 1];
                        }
                    }
                }
            }
    }
    if (!C.stats) return;
    __syncthreads();   // `red` lives in the input tile: every warp must be done reading it
    // pass A: tile mean per column.  red[(wr*8+g)][local column]: 16 partial rows per column
    float* myred = C.red + (wr * 8 + g) * 128 + wc * 8 * NI + 2 * t;
    if (active) {
#pragma unroll
        for (int ni = 0; ni < NI; ++ni)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                float s = 0.f;
#pragma unroll
                for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                    for (int h = 0; h < 2; ++h)
                        if (wr * 32 + mi * 16 + g + 8 * h < rows_valid) s += acc[mi][ni][2 * h + e];
                myred[ni * 8 + e] = s;
            }
    }
    __syncthreads();
    if (tid < cols) {
        float s = 0.f;
#pragma unroll
        for (int q = 0; q < nred; ++q) s += C.red[q * 128 + tid];
        C.tmean[tid] = s / (float)rows_valid;
    }
    __syncthreads();
    // pass B: tile M2 around the tile mean
    if (active) {
#pragma unroll
        for (int ni = 0; ni < NI; ++ni)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const float m = C.tmean[wc * 8 * NI + ni * 8 + 2 * t + e];
                float s = 0.f;
#pragma unroll
                for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                    for (int h = 0; h < 2; ++h)
                        if (wr * 32 + mi * 16 + g + 8 * h < rows_valid) {
                            const float d = acc[mi][ni][2 * h + e] - m;
                            s = fmaf(d, d, s);
                        }
                myred[ni * 8 + e] = s;
            }
    }
    __syncthreads();
    if (tid < cols) {
        float m2 = 0.f;
#pragma unroll
        for (int q = 0; q < nred; ++q) m2 += C.red[q * 128 + tid];
        // Chan merge of (run_cnt, rmean, rm2) with (rows_valid, tmean, m2)
        const int c = pass * 128 + tid;
        const float na = run_cnt, nb = (float)rows_valid, n = na + nb;
        const float d = C.tmean[tid] - C.rmean[c];
        C.rmean[c] += d * (nb / n);
        C.rm2[c] += m2 + d * d * (na * nb / n);
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------
// forward stage kernel
// ------------------------------------------------------------------------------------------
template <bool EXACT, int TMV>
__global__ void __launch_bounds__(NT, 2) tower_fwd_stage(const __grid_constant__ FwdArgs args) {
    const FwdStage& S = args.st[blockIdx.y];
    extern __shared__ float4 smem4[];
    float* sm = reinterpret_cast<float*>(smem4);
    const int K = S.in.K, N = S.N;
    const MmaPlan gp = make_plan(N);
    constexpr int tm = TMV;
    const FwdSmem L = fwd_smem(K, N, S.in.g.n_tab, gp, tm, S.in.stage == 1);
    float *Ws = sm + L.ws, *As = sm + L.as, *sm_bn = sm + L.bn;
    float *rmean = sm + L.rmean, *rm2 = sm + L.rm2;
    int* sm_idx = reinterpret_cast<int*>(sm + L.idx);
    const int tid = threadIdx.x, lda = L.lda;
    const int KE = S.in.g.n_tab * S.in.g.E;
    const bool stats = S.stat_part != nullptr;
    const DropCtx drop = resolve_drop(S.in.a.drop);

    // weights -> smem, [wrows][lda] with bias in column K and zeros elsewhere in the padding
    // (eight independent loads in flight per thread: the staging is pure latency otherwise)
    for (int base = tid; base < gp.wrows * lda; base += NT * 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = base + u * NT;
            const int n = i / lda, c = i - n * lda;
            v[u] = 0.f;
            if (i < gp.wrows * lda && n < N) {
                if (c < K) v[u] = __ldg(S.W + (size_t)n * K + (S.in.stage == 1 ? gcol_stage1(c, KE, S.in.g.n_num) : c));
                else if (c == K) v[u] = __ldg(S.bias + n);
            }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (base + u * NT < gp.wrows * lda) Ws[base + u * NT] = v[u];
    }
    if (S.in.stage > 1) stage_bn_params(S.in.a, K, sm_bn);
    for (int i = tid; i < gp.wrows + 128; i += NT) { rmean[i] = 0.f; rm2[i] = 0.f; }
    float run_cnt = 0.f;   // identical in every thread
    __syncthreads();

    FwdCtx C;
    C.As = As; C.Ws = Ws; C.red = sm + L.red; C.tmean = sm + L.tmean; C.rmean = rmean; C.rm2 = rm2; C.hout = S.hout;
    C.lda = lda; C.k8n = ceil8(K + 1) >> 3; C.N = N; C.stats = stats; C.wcn = tm == 64 ? 4 : 8;
    const long long ntiles = (args.B + tm - 1) / tm;

    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long row0 = tile * tm;
        const int rows_valid = (int)min((long long)tm, args.B - row0);
        build_input_tile(S.in, drop, row0, rows_valid, tm, As, lda, nullptr, 0, sm_bn, sm_idx, args.err);
        __syncthreads();
        for (int pass = 0; pass < gp.passes; ++pass) {
            const int cols = pass_cols(N, pass);
            constexpr int WCN = tm == 64 ? 4 : 8;
            switch (pass_ni(cols, WCN)) {
                case 1: fwd_pass<1, EXACT, WCN>(C, pass, cols, row0, rows_valid, run_cnt); break;
                case 2: fwd_pass<2, EXACT, WCN>(C, pass, cols, row0, rows_valid, run_cnt); break;
                default: if (WCN == 4) fwd_pass<4, EXACT, WCN>(C, pass, cols, row0, rows_valid, run_cnt); break;
            }
        }
        run_cnt += (float)rows_valid;
        __syncthreads();   // As is rebuilt next iteration
    }
    if (stats) {
        float* P = S.stat_part + (size_t)blockIdx.x * (2 * N + 4);
        for (int c = tid; c < N; c += NT) { P[c] = rmean[c]; P[N + c] = rm2[c]; }
        if (tid == 0) P[2 * N] = run_cnt;
    }
}

// merge per-CTA (count, mean, M2) partials in CTA order; write batch mean / istd, update running stats
struct BnFwdFin {
    const float* part; int nparts; int N;
    float* stat;            // [4,N]: mean, istd, (c1, c2 written by the backward finalize)
    float *rm, *rv; long long* nbt;
};
struct BnFwdFinArgs { BnFwdFin t[2]; };
__global__ void bn_fwd_finalize(const __grid_constant__ BnFwdFinArgs args) {
    // one warp per column: lanes Chan-merge a strided subset of the CTA partials, then a fixed shuffle tree
    const BnFwdFin& F = args.t[blockIdx.y];
    const int N = F.N, lane = threadIdx.x & 31;
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c < N) {
        float n = 0.f, mean = 0.f, m2 = 0.f;
        // the partial reads are strided and latency-bound: issue a batch of them before the dependent merges
        constexpr int PF = 4;
        for (int p0 = lane; p0 < F.nparts; p0 += 32 * PF) {
            float nbv[PF], mbv[PF], qbv[PF];
#pragma unroll
            for (int u = 0; u < PF; ++u) {
                const int p = p0 + 32 * u;
                nbv[u] = 0.f; mbv[u] = 0.f; qbv[u] = 0.f;
                if (p < F.nparts) {
                    const float* P = F.part + (size_t)p * (2 * N + 4);
                    nbv[u] = P[2 * N]; mbv[u] = P[c]; qbv[u] = P[N + c];
                }
            }
#pragma unroll
            for (int u = 0; u < PF; ++u) {
                const float nb = nbv[u];
                if (nb <= 0.f) continue;
                float d = mbv[u] - mean, nn = n + nb;
                mean += d * (nb / nn);
                m2 += qbv[u] + d * d * (n * nb / nn);
                n = nn;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float nb = __shfl_down_sync(FULL, n, o), mb = __shfl_down_sync(FULL, mean, o);
            float qb = __shfl_down_sync(FULL, m2, o);
            if (nb > 0.f) {
                float d = mb - mean, nn = n + nb;
                mean += d * (nb / nn);
                m2 += qb + d * d * (n * nb / nn);
                n = nn;
            }
        }
        if (lane == 0) {
            float var_b = m2 / n;
            F.stat[c] = mean;
            F.stat[N + c] = rsqrtf(var_b + BN_EPS);
            if (F.rm) {
                float var_u = n > 1.f ? m2 / (n - 1.f) : var_b;
                F.rm[c] = 0.9f * F.rm[c] + 0.1f * mean;       // momentum 0.1 (nn.BatchNorm1d default)
                F.rv[c] = 0.9f * F.rv[c] + 0.1f * var_u;
            }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && F.nbt && N > 0) *F.nbt += 1;
}

struct BwdCtx {
    const float *As, *Gs, *Wt, *Xh;
    float *red, *rs;
    int lda, ldg, ldxh, n8n, rs_stride, K, KE, n_num;
    float inv_keep;
    bool stage1;
    int wcn;              // column groups of warps, see FwdCtx
};

// one column pass (<= 128 columns of K) of the dX product of a backward tile (+ ReLU/dropout mask, BN-backward sums)
template <int NI, bool EXACT, int WCN>
__device__ __forceinline__ void bwd_dx_pass(const BwdStage& S, const BwdCtx& C, int pass, int cols, long long row0,
                                            int rows_valid) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wr = warp / WCN, wc = warp % WCN, g = lane >> 2, t = lane & 3;
    constexpr int nred = 8 * (8 / WCN);
    const int k0 = pass * 128 + wc * 8 * NI;
    const bool active = wc * 8 * NI < cols;
    float acc[2][NI][4];
    float s1[NI][2], s2[NI][2];
#pragma unroll
    for (int ni = 0; ni < NI; ++ni) { s1[ni][0] = s1[ni][1] = s2[ni][0] = s2[ni][1] = 0.f; }
    if (active) {
        mma_rows<NI, EXACT>(C.Gs, C.ldg, C.Wt, C.ldg, C.n8n, wr * 32, k0, lane, acc);
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int r = wr * 32 + mi * 16 + g + 8 * h;
                if (r < rows_valid) {
#pragma unroll
                    for (int ni = 0; ni < NI; ++ni) {
                        // the two columns a thread owns are adjacent: one 8-byte store each, so a warp's store
                        // instruction fills whole 32-byte sectors (8 rows x 4 lanes x 8 B)
                        const int c0 = k0 + ni * 8 + 2 * t;
                        float v0 = acc[mi][ni][2 * h], v1 = acc[mi][ni][2 * h + 1];
                        if (C.stage1) {
                            if (c0 + 1 < C.KE && !(C.KE & 1)) {
                                if (S.dx_emb)
                                    *reinterpret_cast<float2*>(S.dx_emb + (size_t)(row0 + r) * C.KE + c0) = make_float2(v0, v1);
                            } else {
#pragma unroll
                                for (int e = 0; e < 2; ++e) {
                                    const int c = c0 + e;
                                    const float v = e ? v1 : v0;
                                    if (c >= C.K) continue;
                                    if (c < C.KE) { if (S.dx_emb) S.dx_emb[(size_t)(row0 + r) * C.KE + c] = v; }
                                    else if (S.dx_num) S.dx_num[(size_t)(row0 + r) * C.n_num + (c - C.KE)] = v;
                                }
                            }
                        } else {
                            float dy[2] = {0.f, 0.f};
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const int c = c0 + e;
                                if (c < C.K) {
                                    dy[e] = C.As[r * C.lda + c] > 0.f ? (e ? v1 : v0) * C.inv_keep : 0.f;
                                    if (S.a_bn) { s1[ni][e] += dy[e]; s2[ni][e] = fmaf(dy[e], C.Xh[r * C.ldxh + c], s2[ni][e]); }
                                }
                            }
                            float* dst = S.dy_out + (size_t)(row0 + r) * C.K + c0;
                            if (c0 + 1 < C.K && !(C.K & 1)) *reinterpret_cast<float2*>(dst) = make_float2(dy[0], dy[1]);
                            else { if (c0 < C.K) dst[0] = dy[0]; if (c0 + 1 < C.K) dst[1] = dy[1]; }
                        }
                    }
                }
            }
    }
    if (!S.a_bn) return;
    float* myred = C.red + (wr * 8 + g) * 128 + wc * 8 * NI + 2 * t;
    if (active) {
#pragma unroll
        for (int ni = 0; ni < NI; ++ni)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                myred[ni * 8 + e] = s1[ni][e];
                myred[nred * 128 + ni * 8 + e] = s2[ni][e];
            }
    }
    __syncthreads();
    if (tid < 2 * cols) {
        const int which = tid >= cols, cl = tid - which * cols;
        float tsum = 0.f;
#pragma unroll
        for (int q = 0; q < nred; ++q) tsum += C.red[which * nred * 128 + q * 128 + cl];
        C.rs[which * C.rs_stride + pass * 128 + cl] += tsum;
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------
// backward stage kernel
// ------------------------------------------------------------------------------------------
constexpr int DW_MAX_TILES = 16;   // m16n8 output tiles of dW per warp (N*(K+1) <= 16384)

template <bool EXACT, int TMV>
__global__ void __launch_bounds__(NT, 2) tower_bwd_stage(const __grid_constant__ BwdArgs args) {
    const BwdStage& S = args.st[blockIdx.y];
    const long long B = args.B;
    extern __shared__ float4 smem4[];
    float* sm = reinterpret_cast<float*>(smem4);
    const int K = S.in.K, N = S.N;
    const MmaPlan gx = make_plan(K);   // dX: [TM, K] output, reduction over N
    constexpr int tm = TMV;
    const BwdSmem L = bwd_smem(K, N, S.in.g.n_tab, S.a_bn, S.need_dx, gx, tm, S.in.stage == 1);
    float *Wt = sm + L.wt, *As = sm + L.as, *Gs = sm + L.gs, *Gt = sm + L.gt, *Xh = S.a_bn ? sm + L.xh : nullptr;
    float *sm_bna = sm + L.bna, *sm_bng = sm + L.bng, *rs = sm + L.rs;
    int* sm_idx = reinterpret_cast<int*>(sm + L.idx);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int lda = L.lda, ldg = L.ldg, ldgt = L.ldgt, ldxh = L.ldxh;
    const int KE = S.in.g.n_tab * S.in.g.E, n_num = S.in.g.n_num;
    const int Np = (N + 3) & ~3;
    const bool stage1 = S.in.stage == 1;
    const DropCtx drop = resolve_drop(S.in.a.drop);
    const int rs_stride = gx.wrows + 128;

    // W^T -> smem: Wt[k'][n] (k' in tile column order), zero padded
    if (S.need_dx) {
        for (int base = tid; base < gx.wrows * ldg; base += NT * 8) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = base + u * NT;
                const int k = i / ldg, n = i - k * ldg;
                v[u] = (i < gx.wrows * ldg && k < K && n < N)
                           ? __ldg(S.W + (size_t)n * K + (stage1 ? gcol_stage1(k, KE, n_num) : k)) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (base + u * NT < gx.wrows * ldg) Wt[base + u * NT] = v[u];
        }
    }
    if (!stage1) stage_bn_params(S.in.a, K, sm_bna);
    if (S.g_mode) {
        for (int c = tid; c < N; c += NT) {
            float sv = S.g_var_or_istd[c];
            float istd = S.g_mode == 2 ? rsqrtf(sv + BN_EPS) : sv;
            sm_bng[c] = S.g_mean[c];
            sm_bng[Np + c] = istd;
            sm_bng[2 * Np + c] = S.g_gamma[c] * istd;
            sm_bng[3 * Np + c] = S.g_mode == 1 ? S.g_c1[c] : 0.f;
            sm_bng[4 * Np + c] = S.g_mode == 1 ? S.g_c2[c] : 0.f;
        }
    }
    for (int i = tid; S.a_bn && i < 2 * rs_stride; i += NT) rs[i] = 0.f;
    // the G^T tile's padding rows (n >= N) stay zero for the whole kernel
    for (int i = tid; i < ((N + 15) & ~15) * ldgt; i += NT) Gt[i] = 0.f;

    // dW = G^T . [A | 1]: m16n8 output tiles (n-tile, k-tile), tile q = warp + 8*i belongs to this warp
    const int NTN = (N + 15) >> 4, KT = ceil8(K + 1) >> 3, n_dw_tiles = NTN * KT;
    float dw[DW_MAX_TILES][4];
#pragma unroll
    for (int i = 0; i < DW_MAX_TILES; ++i) dw[i][0] = dw[i][1] = dw[i][2] = dw[i][3] = 0.f;
    const int g = lane >> 2, t = lane & 3;

    BwdCtx C;
    C.As = As; C.Gs = Gs; C.Wt = Wt; C.Xh = Xh; C.red = sm + L.red; C.rs = rs;
    C.lda = lda; C.ldg = ldg; C.ldxh = ldxh; C.n8n = ceil8(N) >> 3; C.rs_stride = rs_stride;
    C.K = K; C.KE = KE; C.n_num = n_num;
    C.inv_keep = (!stage1 && drop.active) ? drop.inv_keep : 1.f;
    C.stage1 = stage1; C.wcn = tm == 64 ? 4 : 8;
    __syncthreads();

    const long long ntiles = (B + tm - 1) / tm;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long row0 = tile * tm;
        const int rows_valid = (int)min((long long)tm, B - row0);
        build_input_tile(S.in, drop, row0, rows_valid, tm, As, lda, Xh, ldxh, sm_bna, sm_idx, nullptr);
        // incoming-gradient tile G [TM, N] (zero padded to ldg) and its transpose Gt [N, TM];
        // loads of two items (8 gradient + 8 saved-activation values) are issued before either is used
        {
            const int ldg4 = ldg >> 2;
            const bool vecn = (N & 3) == 0;
            for (int base = tid; base < tm * ldg4; base += NT * 2) {
                float dyv[2][4], hsv[2][4];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int i = base + u * NT;
                    const int r = i / ldg4, c4 = i - r * ldg4;
#pragma unroll
                    for (int e = 0; e < 4; ++e) { dyv[u][e] = 0.f; hsv[u][e] = 0.f; }
                    if (i < tm * ldg4 && r < rows_valid && 4 * c4 < N) {
                        const size_t off = (size_t)(row0 + r) * N + 4 * c4;
                        if (vecn) {
                            const float4 d4 = *reinterpret_cast<const float4*>(S.gin + off);
                            dyv[u][0] = d4.x; dyv[u][1] = d4.y; dyv[u][2] = d4.z; dyv[u][3] = d4.w;
                            if (S.g_mode == 1) {
                                const float4 h4 = *reinterpret_cast<const float4*>(S.hs + off);
                                hsv[u][0] = h4.x; hsv[u][1] = h4.y; hsv[u][2] = h4.z; hsv[u][3] = h4.w;
                            }
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                if (4 * c4 + e < N) {
                                    dyv[u][e] = S.gin[off + e];
                                    if (S.g_mode == 1) hsv[u][e] = S.hs[off + e];
                                }
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int i = base + u * NT;
                    if (i >= tm * ldg4) continue;
                    const int r = i / ldg4, c4 = i - r * ldg4;
                    float gv[4] = {0.f, 0.f, 0.f, 0.f};
                    if (r < rows_valid) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int c = 4 * c4 + e;
                            if (c < N) {
                                float dy = dyv[u][e];
                                if (S.g_mode == 1) {
                                    const float xh = (hsv[u][e] - sm_bng[c]) * sm_bng[Np + c];
                                    dy = sm_bng[2 * Np + c] * (dy - sm_bng[3 * Np + c] - xh * sm_bng[4 * Np + c]);
                                } else if (S.g_mode == 2) {
                                    dy *= sm_bng[2 * Np + c];
                                }
                                gv[e] = dy;
                            }
                        }
                    }
                    *reinterpret_cast<float4*>(Gs + r * ldg + 4 * c4) = make_float4(gv[0], gv[1], gv[2], gv[3]);
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (4 * c4 + e < N) Gt[(4 * c4 + e) * ldgt + r] = gv[e];
                }
            }
        }
        __syncthreads();

        // ---- dW accumulation over the 64 rows of this tile (8 k-steps of 8 rows) ----
#pragma unroll
        for (int i = 0; i < DW_MAX_TILES; ++i) {
            const int q = warp + 8 * i;
            if (q < n_dw_tiles) {
                const int nt = q % NTN, kt = q / NTN;
                const int mat = lane >> 3, rin = lane & 7;
                const float* gp_ = Gt + (nt * 16 + rin + (mat & 1) * 8) * ldgt + (mat >> 1) * 4;   // A: (m = n, k = row)
                const float* ap_ = As + t * lda + kt * 8 + g;              // B fragment: (k = row, n = k_in)
#pragma unroll 2
                for (int r8 = 0; r8 < tm / 8; ++r8) {
                    uint32_t raw[4], ah[4], al[4], bh[2], bl[2];
                    ldmatrix_x4(gp_ + r8 * 8, raw);
#pragma unroll
                    for (int e = 0; e < 4; ++e) split_frag<EXACT>(raw[e], ah[e], al[e]);
                    const float* aq = ap_ + r8 * 8 * lda;
                    split_frag<EXACT>(__float_as_uint(aq[0]), bh[0], bl[0]);
                    split_frag<EXACT>(__float_as_uint(aq[4 * lda]), bh[1], bl[1]);
                    mma_acc<EXACT>(dw[i], ah, al, bh, bl);
                }
            }
        }

        if (S.need_dx) {
            for (int pass = 0; pass < gx.passes; ++pass) {
                const int cols = pass_cols(K, pass);
                constexpr int WCN = tm == 64 ? 4 : 8;
                switch (pass_ni(cols, WCN)) {
                    case 1: bwd_dx_pass<1, EXACT, WCN>(S, C, pass, cols, row0, rows_valid); break;
                    case 2: bwd_dx_pass<2, EXACT, WCN>(S, C, pass, cols, row0, rows_valid); break;
                    default: if (WCN == 4) bwd_dx_pass<4, EXACT, WCN>(S, C, pass, cols, row0, rows_valid); break;
                }
            }
        }
        __syncthreads();   // tiles are rebuilt next iteration
    }

    // per-CTA partial outputs: dW[n][k] (column K = bias gradient), global column order
    {
        float* P = S.dW_part + (size_t)blockIdx.x * N * (K + 1);
#pragma unroll
        for (int i = 0; i < DW_MAX_TILES; ++i) {
            const int q = warp + 8 * i;
            if (q < n_dw_tiles) {
                const int nt = q % NTN, kt = q / NTN;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int n = nt * 16 + g + 8 * (c >> 1), k = kt * 8 + 2 * t + (c & 1);
                    if (n < N && k <= K) {
                        const int kgl = (k < K && stage1) ? gcol_stage1(k, KE, n_num) : k;
                        P[(size_t)n * (K + 1) + kgl] = dw[i][c];
                    }
                }
            }
        }
    }
    if (S.a_bn && S.need_dx) {
        float* P = S.sum_part + (size_t)blockIdx.x * 2 * K;
        for (int c = tid; c < K; c += NT) { P[c] = rs[c]; P[K + c] = rs[rs_stride + c]; }
    }
}

// dW[n][k] = sum over CTA partials (fixed order), bias gradient from the extra column
struct ReduceW { const float* part; int nparts; int N, K; float* dW; float* db; };
struct ReduceWArgs { ReduceW t[2]; };
__global__ void __launch_bounds__(256) reduce_dw(const __grid_constant__ ReduceWArgs args) {
    // 32 consecutive elements per CTA x 8 partial lanes; lane pl sums partials pl, pl+8, ... then the 8 lane
    // sums are added in lane order: fixed association, fully coalesced 128-byte reads
    const ReduceW& R = args.t[blockIdx.y];
    __shared__ float sm[8][33];
    const int stride = R.N * (R.K + 1);
    const int el = threadIdx.x & 31, pl = threadIdx.x >> 5;
    for (int base = blockIdx.x * 32; base < stride; base += gridDim.x * 32) {
        const int i = base + el;
        float s = 0.f;
        if (i < stride) {
#pragma unroll 4
            for (int p = pl; p < R.nparts; p += 8) s += R.part[(size_t)p * stride + i];
        }
        sm[pl][el] = s;
        __syncthreads();
        if (pl == 0 && i < stride) {
            float tot = 0.f;
#pragma unroll
            for (int q = 0; q < 8; ++q) tot += sm[q][el];
            int n = i / (R.K + 1), k = i - n * (R.K + 1);
            if (k < R.K) R.dW[(size_t)n * R.K + k] = tot;
            else if (R.db) R.db[n] = tot;
        }
        __syncthreads();
    }
}

// sums of dy and dy*xhat over the batch -> BN affine grads and the (c1, c2) coefficients of the BN backward
struct BnBwdFin { const float* part; int nparts; int K; float B; float* dgamma; float* dbeta; float* stat; };
struct BnBwdFinArgs { BnBwdFin t[2]; };
__global__ void bn_bwd_finalize(const __grid_constant__ BnBwdFinArgs args) {
    const BnBwdFin& F = args.t[blockIdx.y];
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);   // one warp per column
    if (c >= F.K) return;
    float s1 = 0.f, s2 = 0.f;
    for (int p = lane; p < F.nparts; p += 32) {
        s1 += F.part[(size_t)p * 2 * F.K + c];
        s2 += F.part[(size_t)p * 2 * F.K + F.K + c];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_down_sync(FULL, s1, o);
        s2 += __shfl_down_sync(FULL, s2, o);
    }
    if (lane == 0) {
        if (F.dbeta) F.dbeta[c] = s1;
        if (F.dgamma) F.dgamma[c] = s2;
        F.stat[2 * F.K + c] = s1 / F.B;
        F.stat[3 * F.K + c] = s2 / F.B;
    }
}

__global__ void counter_advance_kernel(unsigned long long* c, unsigned long long inc) { *c += inc; }

__global__ void dropout_mask_kernel(uint8_t* mask, long long B, int width, DropCtx d) {
    const int W4 = (width + 3) >> 2;
    long long total = B * W4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long r = i / W4;
        int c4 = (int)(i - r * W4);
        Philox4 w = drop_words(d, r, c4);
        for (int e = 0; e < 4; ++e) {
            int c = 4 * c4 + e;
            if (c < width) mask[r * width + c] = d.active ? (drop_keep(d, w, e) ? 1 : 0) : 1;
        }
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static int validate_tower(const cfm_tower_t& t) {
    CFM_REQUIRE(t.n_tables >= 0 && t.n_tables <= CFM_MAX_TABLES, CFM_ERR_INVALID, "n_tables %lld outside [0,%d]",
                (long long)t.n_tables, CFM_MAX_TABLES);
    CFM_REQUIRE(t.n_num >= 0 && t.emb_dim >= 0 && t.h1 > 0 && t.h2 > 0 && t.d_out > 0, CFM_ERR_INVALID,
                "bad tower dims");
    CFM_REQUIRE(t.n_num + t.n_tables * t.emb_dim > 0, CFM_ERR_INVALID, "tower has no inputs");
    CFM_REQUIRE(t.w1 && t.b1 && t.w2 && t.b2 && t.w3 && t.b3 && t.bn1_w && t.bn1_b && t.bn1_rm && t.bn1_rv,
                CFM_ERR_INVALID, "null tower parameter");
    CFM_REQUIRE(!t.bn2 || (t.bn2_w && t.bn2_b && t.bn2_rm && t.bn2_rv), CFM_ERR_INVALID, "null bn2 parameter");
    CFM_REQUIRE(t.h1_raw && t.h2_raw && t.out && t.bn1_stat && t.scratch, CFM_ERR_INVALID, "null tower workspace");
    CFM_REQUIRE(!t.bn2 || t.bn2_stat, CFM_ERR_INVALID, "null bn2_stat");
    CFM_REQUIRE(t.drop1 >= 0 && t.drop1 < 1 && t.drop2 >= 0 && t.drop2 < 1, CFM_ERR_INVALID, "dropout p outside [0,1)");
    CFM_REQUIRE(t.precision == 0 || t.precision == 1, CFM_ERR_INVALID, "precision must be 0 (fp32) or 1 (tf32)");
    for (int i = 0; i < t.n_tables; ++i)
        CFM_REQUIRE(t.tables[i] && ((uintptr_t)t.tables[i] & 15) == 0, CFM_ERR_INVALID,
                    "table %d null or not 16-byte aligned", i);
    return CFM_OK;
}

static void fill_gather(GatherSrc& g, const cfm_tower_t& t) {
    g.n_num = (int)t.n_num; g.n_tab = (int)t.n_tables; g.E = (int)t.emb_dim;
    g.x_num = t.x_num; g.x_cat = (const long long*)t.x_cat;
    for (int i = 0; i < CFM_MAX_TABLES; ++i) {
        g.tab[i] = i < t.n_tables ? t.tables[i] : nullptr;
        g.tab_rows[i] = i < t.n_tables ? t.table_rows[i] : 0;
    }
}

// input description of stage s (1..3) of tower t
static InputDesc input_desc(const cfm_tower_t& t, int s, bool training, uint64_t seed, uint64_t offset,
                            const uint64_t* dev_off) {
    InputDesc in{};
    in.stage = s;
    fill_gather(in.g, t);
    if (s == 1) {
        in.K = (int)(t.n_num + t.n_tables * t.emb_dim);
        in.a.drop = make_drop(0.0, false, 0, 0, 0, 0);
    } else if (s == 2) {
        in.K = (int)t.h1;
        in.a.h = t.h1_raw;
        in.a.bn_mode = training ? 1 : 2;
        in.a.mean = training ? t.bn1_stat : t.bn1_rm;
        in.a.var_or_istd = training ? t.bn1_stat + t.h1 : t.bn1_rv;
        in.a.gamma = t.bn1_w; in.a.beta = t.bn1_b;
        in.a.drop = make_drop(t.drop1, training, seed, offset, t.tower_id, 0);
    } else {
        in.K = (int)t.h2;
        in.a.h = t.h2_raw;
        in.a.bn_mode = t.bn2 ? (training ? 1 : 2) : 0;
        if (t.bn2) {
            in.a.mean = training ? t.bn2_stat : t.bn2_rm;
            in.a.var_or_istd = training ? t.bn2_stat + t.h2 : t.bn2_rv;
            in.a.gamma = t.bn2_w; in.a.beta = t.bn2_b;
        }
        in.a.drop = make_drop(t.drop2, training, seed, offset, t.tower_id, 1);
    }
    in.a.drop.dev_off = (const unsigned long long*)dev_off;
    return in;
}

static int stage_N(const cfm_tower_t& t, int s) { return (int)(s == 1 ? t.h1 : s == 2 ? t.h2 : t.d_out); }
static int stage_K(const cfm_tower_t& t, int s) {
    return (int)(s == 1 ? t.n_num + t.n_tables * t.emb_dim : s == 2 ? t.h1 : t.h2);
}

// persistent CTAs per tower: one per SM, two for 32-row tiles (cfm_device_info reports the maximum: scratch sizing)
static int tower_ctas(int tm) { return tm == 32 ? 2 * sm_count() : sm_count(); }
// 32-row tiles exactly when they turn a one-CTA-per-SM stage (weights + tile > half the SM's shared memory) into a
// two-CTA one: the stage is latency-bound, a second resident CTA is worth more than the larger tile
static int pick_tm(size_t smem64, size_t smem32) {
    const size_t two_per_sm = 113 * 1024;
    return (smem64 > two_per_sm && smem32 <= two_per_sm) ? 32 : 64;
}

static size_t fwd_smem_bytes(const cfm_tower_t& t, int s, int tm) {
    MmaPlan gp = make_plan(stage_N(t, s));
    return (size_t)fwd_smem(stage_K(t, s), stage_N(t, s), (int)t.n_tables, gp, tm, s == 1).total_floats * 4;
}
static size_t bwd_smem_bytes(const cfm_tower_t& t, int s, int a_bn, int need_dx, int tm) {
    MmaPlan gx = make_plan(stage_K(t, s));
    return (size_t)bwd_smem(stage_K(t, s), stage_N(t, s), (int)t.n_tables, a_bn, need_dx, gx, tm, s == 1).total_floats * 4;
}

}  // namespace cfm

using namespace cfm;

// Per-CTA floats of a tower's scratch: one weight-gradient partial region per layer (so that the reduction of layer s
// can run beside the backward of layer s - 1) followed by the small BatchNorm partials (forward statistics, backward
// sums).  Regions are laid out for the maximum CTA count cfm_device_info reports.
static int64_t dw_region_floats(const cfm_tower_t& t, int s) { return (int64_t)stage_N(t, s) * (stage_K(t, s) + 1); }
static int64_t max_ctas() { return 2 * (int64_t)sm_count(); }
static float* dw_region(const cfm_tower_t& t, int s) {
    int64_t off = 0;
    for (int q = 3; q > s; --q) off += dw_region_floats(t, q);
    return t.scratch + max_ctas() * off;
}
static float* misc_region(const cfm_tower_t& t) {
    return t.scratch + max_ctas() * (dw_region_floats(t, 1) + dw_region_floats(t, 2) + dw_region_floats(t, 3));
}
extern "C" int64_t cfm_tower_scratch_floats(const cfm_tower_t* t) {
    int64_t m = 0, d = 0;
    for (int s = 1; s <= 3; ++s) {
        int64_t n = stage_N(*t, s), k = stage_K(*t, s);
        m += n * (k + 1);
        d = n > d ? n : d;
        d = k > d ? k : d;
    }
    return m + 4 * d + 16;
}

// Side stream of the library (one per device): work that is off the critical path of a step (the fixed-order
// reduction of the per-CTA weight-gradient partials) is forked onto it and joined at the end of the call; under
// stream capture the fork / join events become graph edges.
struct SideStream {
    cudaStream_t stream = nullptr;
    cudaEvent_t fork[4] = {nullptr, nullptr, nullptr, nullptr}, join = nullptr;
};
static SideStream* side_stream() {
    static SideStream side[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    SideStream& S = side[dev];
    if (!S.stream) {
        if (cudaStreamCreateWithFlags(&S.stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        for (int i = 0; i < 4; ++i) cudaEventCreateWithFlags(&S.fork[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&S.join, cudaEventDisableTiming);
    }
    return &S;
}

extern "C" int cfm_towers_fwd(const cfm_tower_t* towers, int64_t n_towers, int64_t B, int64_t training,
                              uint64_t seed, uint64_t offset, const uint64_t* rng_offset_dev, int32_t* err_flag,
                              void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CFM_REQUIRE(towers && n_towers >= 1 && n_towers <= 2, CFM_ERR_INVALID, "n_towers must be 1 or 2");
    CFM_REQUIRE(B >= 1, CFM_ERR_INVALID, "B must be >= 1 (got %lld)", (long long)B);
    CFM_REQUIRE(!(training && B == 1), CFM_ERR_BATCHNORM_B1,
                "Expected more than 1 value per channel when training (B=1)");
    for (int i = 0; i < n_towers; ++i) { int rc = validate_tower(towers[i]); if (rc) return rc; }
    static bool attr_set = false;
    if (!attr_set) {
        CFM_CHECK_CUDA(cudaFuncSetAttribute(tower_fwd_stage<true, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_SMEM));
        CFM_CHECK_CUDA(cudaFuncSetAttribute(tower_fwd_stage<false, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_SMEM));
        attr_set = true;
    }
    const bool exact = towers[0].precision == 0;
    // layers the tcgen05 stage kernels cover (all towers of the call must qualify); the rest run on mma.sync
    bool use_tc[4] = {false, false, false, false};
    for (int s = 1; s <= 3; ++s) {
        use_tc[s] = true;
        for (int i = 0; i < n_towers; ++i) use_tc[s] = use_tc[s] && tc_fwd_supported(towers[i], s);
    }
    if (towers[0].wimg) { int rc = tc_prep_launch(towers, (int)n_towers, stream); if (rc) return rc; }
    for (int s = 1; s <= 3; ++s) {
        FwdArgs a{};
        a.B = B; a.err = err_flag; a.exact = exact ? 1 : 0;
        size_t smem = 0, smem32 = 0;
        bool any_stats = false;
        for (int i = 0; i < n_towers; ++i) {
            smem = std::max(smem, fwd_smem_bytes(towers[i], s, 64));
            smem32 = std::max(smem32, fwd_smem_bytes(towers[i], s, 32));
        }
        // forward stays at 64-row tiles: with 32 rows every warp would reload the whole A tile for one n-tile of
        // output (measured: no gain), and the statistics partials would double
        (void)smem32;
        a.tm = 64;
        const long long ntiles = (B + a.tm - 1) / a.tm;
        // a stage whose CTAs fit exactly twice per SM gets two persistent CTAs per SM and tower, so that the SM keeps
        // 16 warps while one tower's CTAs run (both towers' grids are otherwise resident one after the other)
        const bool two = smem > 76 * 1024 && smem <= 113 * 1024;
        int ctas = (int)std::min<long long>(ntiles, two ? tower_ctas(32) : tower_ctas(64));
        int ctas_t[2] = {ctas, ctas};                    // per-tower CTA (= partial) counts
        for (int i = 0; i < n_towers; ++i) {
            const cfm_tower_t& t = towers[i];
            FwdStage& S = a.st[i];
            S.in = input_desc(t, s, training != 0, seed, offset, rng_offset_dev);
            S.N = stage_N(t, s);
            S.W = s == 1 ? t.w1 : s == 2 ? t.w2 : t.w3;
            S.bias = s == 1 ? t.b1 : s == 2 ? t.b2 : t.b3;
            S.hout = s == 1 ? t.h1_raw : s == 2 ? t.h2_raw : t.out;
            bool bn_after = s == 1 || (s == 2 && t.bn2);
            S.stat_part = (training && bn_after) ? misc_region(t) : nullptr;
            any_stats |= S.stat_part != nullptr;
            size_t need = fwd_smem_bytes(t, s, a.tm);
            CFM_REQUIRE(use_tc[s] || need <= (size_t)MAX_SMEM, CFM_ERR_UNSUPPORTED,
                        "tower stage %d needs %zu B shared memory (> %d): layer %dx%d too large", s, need, MAX_SMEM,
                        S.N, S.in.K);
        }
        if (use_tc[s]) {
            ProfScope prof(PROF_FWD1 + s - 1, stream);
            int rc = tc_fwd_launch(a, towers, (int)n_towers, s, ctas_t, stream);   // the towers share the SMs side by side
            if (rc) return rc;
        } else {
            ProfScope prof(PROF_FWD1 + s - 1, stream);
            if (exact) tower_fwd_stage<true, 64><<<dim3(ctas, (unsigned)n_towers), NT, smem, stream>>>(a);
            else tower_fwd_stage<false, 64><<<dim3(ctas, (unsigned)n_towers), NT, smem, stream>>>(a);
            CFM_LAUNCH_CHECK();
        }
        if (any_stats) {
            BnFwdFinArgs f{};
            for (int i = 0; i < n_towers; ++i) {
                const cfm_tower_t& t = towers[i];
                bool bn_after = s == 1 || (s == 2 && t.bn2);
                BnFwdFin& F = f.t[i];
                F.part = misc_region(t); F.nparts = ctas_t[i]; F.N = bn_after ? stage_N(t, s) : 0;
                F.stat = s == 1 ? t.bn1_stat : t.bn2_stat;
                F.rm = s == 1 ? t.bn1_rm : t.bn2_rm;
                F.rv = s == 1 ? t.bn1_rv : t.bn2_rv;
                F.nbt = (long long*)(s == 1 ? t.bn1_nbt : t.bn2_nbt);
                if (!bn_after) F.nbt = nullptr;
            }
            {
                ProfScope prof(PROF_REDUCE, stream);
                int maxN = 1; for (int i = 0; i < n_towers; ++i) maxN = std::max(maxN, f.t[i].N);
                bn_fwd_finalize<<<dim3(ceil_div(maxN, 8), (unsigned)n_towers), 256, 0, stream>>>(f);
            }
            CFM_LAUNCH_CHECK();
        }
    }
    return CFM_OK;
}

static int launch_bwd(const BwdArgs& a, int ctas, int n_towers, size_t smem, bool exact, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        CFM_CHECK_CUDA(cudaFuncSetAttribute(tower_bwd_stage<true, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_SMEM));
        CFM_CHECK_CUDA(cudaFuncSetAttribute(tower_bwd_stage<false, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_SMEM));
        CFM_CHECK_CUDA(cudaFuncSetAttribute(tower_bwd_stage<true, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_SMEM));
        CFM_CHECK_CUDA(cudaFuncSetAttribute(tower_bwd_stage<false, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_SMEM));
        attr_set = true;
    }
    const dim3 grid(ctas, (unsigned)n_towers);
    if (a.tm == 32) {
        if (exact) tower_bwd_stage<true, 32><<<grid, NT, smem, stream>>>(a);
        else tower_bwd_stage<false, 32><<<grid, NT, smem, stream>>>(a);
    } else {
        if (exact) tower_bwd_stage<true, 64><<<grid, NT, smem, stream>>>(a);
        else tower_bwd_stage<false, 64><<<grid, NT, smem, stream>>>(a);
    }
    CFM_LAUNCH_CHECK();
    return CFM_OK;
}

extern "C" int cfm_towers_bwd(const cfm_tower_t* towers, const cfm_tower_grads_t* grads, int64_t n_towers, int64_t B,
                              int64_t training, uint64_t seed, uint64_t offset, const uint64_t* rng_offset_dev,
                              void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CFM_REQUIRE(towers && grads && n_towers >= 1 && n_towers <= 2, CFM_ERR_INVALID, "n_towers must be 1 or 2");
    CFM_REQUIRE(B >= 1, CFM_ERR_INVALID, "B must be >= 1");
    for (int i = 0; i < n_towers; ++i) {
        int rc = validate_tower(towers[i]);
        if (rc) return rc;
        const cfm_tower_grads_t& g = grads[i];
        CFM_REQUIRE(g.g_out && g.dw1 && g.db1 && g.dw2 && g.db2 && g.dw3 && g.db3 && g.dbn1_w && g.dbn1_b && g.dy1 &&
                        g.dy2, CFM_ERR_INVALID, "null gradient buffer");
        CFM_REQUIRE(!towers[i].bn2 || (g.dbn2_w && g.dbn2_b), CFM_ERR_INVALID, "null bn2 gradient buffer");
    }
    SideStream* side = side_stream();
    for (int s = 3; s >= 1; --s) {
        BwdArgs a{};
        a.B = B;
        size_t smem = 0, smem32 = 0;
        int max_groups = 0;
        bool any_sums = false;
        for (int i = 0; i < n_towers; ++i) {
            const cfm_tower_t& t = towers[i];
            const bool a_bn_i = s >= 2 && (s == 2 || t.bn2);         // BN precedes this stage's input activation
            const bool need_dx_i = s > 1 || grads[i].dx_emb || grads[i].dx_num;
            smem = std::max(smem, bwd_smem_bytes(t, s, a_bn_i, need_dx_i, 64));
            smem32 = std::max(smem32, bwd_smem_bytes(t, s, a_bn_i, need_dx_i, 32));
        }
        a.tm = pick_tm(smem, smem32);
        a.exact = towers[0].precision == 0 ? 1 : 0;
        bool use_tc = true;                               // tcgen05 stage kernel when every tower of the call qualifies
        for (int i = 0; i < n_towers; ++i) {
            const cfm_tower_t& t = towers[i];
            const bool a_bn_i = s >= 2 && (s == 2 || t.bn2);
            const bool need_dx_i = s > 1 || grads[i].dx_emb || grads[i].dx_num;
            use_tc = use_tc && tc_bwd_supported(t, s, a_bn_i, need_dx_i);
        }
        if (use_tc) a.tm = 64;
        const long long ntiles = (B + a.tm - 1) / a.tm;
        int ctas = (int)std::min<long long>(ntiles, use_tc ? sm_count() : tower_ctas(a.tm));
        int ctas_t[2] = {ctas, ctas};                    // per-tower CTA (= partial) counts
        smem = 0;
        for (int i = 0; i < n_towers; ++i) {
            const cfm_tower_t& t = towers[i];
            const cfm_tower_grads_t& g = grads[i];
            BwdStage& S = a.st[i];
            S.in = input_desc(t, s, training != 0, seed, offset, rng_offset_dev);
            S.N = stage_N(t, s);
            S.W = s == 1 ? t.w1 : s == 2 ? t.w2 : t.w3;
            S.a_bn = s >= 2 && S.in.a.bn_mode != 0;
            S.gin = s == 3 ? g.g_out : s == 2 ? g.dy2 : g.dy1;
            bool bn_s = s == 1 || (s == 2 && t.bn2);
            S.g_mode = (s < 3 && bn_s) ? (training ? 1 : 2) : 0;
            if (S.g_mode) {
                const float* stat = s == 1 ? t.bn1_stat : t.bn2_stat;
                int h = stage_N(t, s);
                S.hs = s == 1 ? t.h1_raw : t.h2_raw;
                S.g_mean = training ? stat : (s == 1 ? t.bn1_rm : t.bn2_rm);
                S.g_var_or_istd = training ? stat + h : (s == 1 ? t.bn1_rv : t.bn2_rv);
                S.g_gamma = s == 1 ? t.bn1_w : t.bn2_w;
                S.g_c1 = stat + 2 * h; S.g_c2 = stat + 3 * h;
            }
            S.dW_part = dw_region(t, s);
            S.dy_out = s == 3 ? g.dy2 : s == 2 ? g.dy1 : nullptr;
            S.sum_part = misc_region(t);
            S.dx_emb = s == 1 ? g.dx_emb : nullptr;
            S.dx_num = s == 1 ? g.dx_num : nullptr;
            S.need_dx = s > 1 || g.dx_emb || g.dx_num;
            any_sums |= S.a_bn != 0;
            size_t need = use_tc ? 0 : bwd_smem_bytes(t, s, S.a_bn, S.need_dx, a.tm);
            CFM_REQUIRE(need <= (size_t)MAX_SMEM, CFM_ERR_UNSUPPORTED,
                        "tower bwd stage %d needs %zu B shared memory (> %d)", s, need, MAX_SMEM);
            smem = std::max(smem, need);
            int groups = ceil_div(S.N, 4) * ceil_div(S.in.K + 1, 4);
            max_groups = std::max(max_groups, groups);
        }
        // scratch must hold ctas * (N*(K+1) + 2K) floats: guaranteed by cfm_tower_scratch_floats
        for (int i = 0; i < n_towers && !use_tc; ++i) {
            const BwdStage& S = a.st[i];
            CFM_REQUIRE(ceil_div(S.N, 16) * (ceil8(S.in.K + 1) >> 3) <= 8 * DW_MAX_TILES, CFM_ERR_UNSUPPORTED,
                        "tower bwd stage %d: layer %dx%d too large for the per-warp dW tiles", s, S.N, S.in.K);
        }
        int rc;
        {
            ProfScope prof(PROF_BWD1 + s - 1, stream);
            if (use_tc) rc = tc_bwd_launch(a, towers, (int)n_towers, s, ctas_t, stream);      // the towers share the SMs side by side
            else rc = launch_bwd(a, ctas, (int)n_towers, smem, towers[0].precision == 0, stream);
        }
        if (rc) return rc;
        (void)max_groups;
        // fixed-order reduction of the per-CTA weight-gradient partials: nothing downstream in this call needs it,
        // so it runs on the side stream while the next layer's backward proceeds
        {
            ReduceWArgs r{};
            for (int i = 0; i < n_towers; ++i) {
                const cfm_tower_t& t = towers[i];
                const cfm_tower_grads_t& g = grads[i];
                ReduceW& R = r.t[i];
                R.part = dw_region(t, s); R.nparts = ctas_t[i]; R.N = a.st[i].N; R.K = a.st[i].in.K;
                R.dW = s == 1 ? g.dw1 : s == 2 ? g.dw2 : g.dw3;
                R.db = s == 1 ? g.db1 : s == 2 ? g.db2 : g.db3;
            }
            int maxS = 1; for (int i = 0; i < n_towers; ++i) maxS = std::max(maxS, r.t[i].N * (r.t[i].K + 1));
            cudaStream_t rs = stream;
            if (side) {
                CFM_CHECK_CUDA(cudaEventRecord(side->fork[s], stream));
                CFM_CHECK_CUDA(cudaStreamWaitEvent(side->stream, side->fork[s], 0));
                rs = side->stream;
            }
            ProfScope prof_red(PROF_REDUCE, rs);
            reduce_dw<<<dim3(std::min(ceil_div(maxS, 32), 592), (unsigned)n_towers), 256, 0, rs>>>(r);
            CFM_LAUNCH_CHECK();
        }
        ProfScope prof_red(PROF_REDUCE, stream);
        if (any_sums) {
            BnBwdFinArgs f{};
            for (int i = 0; i < n_towers; ++i) {
                const cfm_tower_t& t = towers[i];
                const cfm_tower_grads_t& g = grads[i];
                const BwdStage& S = a.st[i];
                BnBwdFin& F = f.t[i];
                F.part = S.sum_part; F.nparts = ctas_t[i]; F.K = S.a_bn ? S.in.K : 0; F.B = (float)B;
                F.dgamma = s == 3 ? g.dbn2_w : g.dbn1_w;
                F.dbeta = s == 3 ? g.dbn2_b : g.dbn1_b;
                F.stat = s == 3 ? t.bn2_stat : t.bn1_stat;
            }
            int maxK = 1; for (int i = 0; i < n_towers; ++i) maxK = std::max(maxK, f.t[i].K);
            bn_bwd_finalize<<<dim3(ceil_div(maxK, 8), (unsigned)n_towers), 256, 0, stream>>>(f);
            CFM_LAUNCH_CHECK();
        }
    }
    if (side) {
        CFM_CHECK_CUDA(cudaEventRecord(side->join, side->stream));
        CFM_CHECK_CUDA(cudaStreamWaitEvent(stream, side->join, 0));
    }
    return CFM_OK;
}

extern "C" int cfm_dropout_mask(uint8_t* mask, int64_t B, int64_t width, double p, int64_t tower_id, int64_t site,
                                uint64_t seed, uint64_t offset, void* stream_) {
    CFM_REQUIRE(mask && B >= 0 && width > 0, CFM_ERR_INVALID, "bad dropout mask arguments");
    if (B == 0) return CFM_OK;
    DropCtx d = make_drop(p, true, seed, offset, tower_id, (int)site);
    dropout_mask_kernel<<<std::min<long long>(1024, (B * ((width + 3) / 4) + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(
        mask, B, (int)width, d);
    CFM_LAUNCH_CHECK();
    return CFM_OK;
}

extern "C" int cfm_counter_advance(uint64_t* counter_dev, uint64_t inc, void* stream_) {
    CFM_REQUIRE(counter_dev, CFM_ERR_INVALID, "null counter");
    counter_advance_kernel<<<1, 1, 0, (cudaStream_t)stream_>>>((unsigned long long*)counter_dev, inc);
    CFM_LAUNCH_CHECK();
    return CFM_OK;
}
