// Tower stage kernels on the 5th-generation tensor cores: tcgen05.mma kind::tf32, accumulators in TMEM.
//
// replaces ceo_firm_matching/model.py:69-76 and structural_model.py:120-127 (+ their autograd) for layers up to 64
// outputs wide; same stage contract as the mma.sync kernels of tower.cu (one Linear per launch, both towers in one
// grid, identical per-CTA partial formats), so the finalize / reduce kernels and the host orchestration are shared.
//
// Every product is an fp32-class "3xTF32" product: each fp32 operand is split ONCE, by the thread that builds its
// shared-memory image, into hi = tf32(x) and lo = x - hi (tctile.cuh); the tensor core runs a_lo.b_hi + a_hi.b_lo +
// a_hi.b_hi into one TMEM accumulator (precision 1: the last pass only).  All operands are K-major images; where a
// product reduces over the batch rows (weight gradients) the builders write the transposed image.
//
// One persistent CTA per SM and tower, 21 warps with fixed roles:
//   warps 0-3   epilogue : TMEM -> registers (thread <-> accumulator row), bias / masks, global stores, the
//                          cross-row column sums (BatchNorm statistics forward, BatchNorm-backward sums) by
//                          recursive-halving shuffles, accumulated in registers over all tiles of the CTA
//   warp  4     MMA      : one thread issues every tcgen05.mma and the tcgen05.commit that releases a buffer
//   warps 5-20  producers: two groups of eight warps take alternate chunks (32 input columns of a row tile) and
//                          build their operand images in a shared-memory ring: stage 1 gathers embedding rows
//                          (16-byte loads), later stages re-apply BatchNorm / ReLU / dropout to the saved
//                          pre-activations; memory-level parallelism comes from the sixteen warps
// Rings and accumulators are handed over with mbarriers (full / empty pairs); nothing in the steady state uses a
// CTA-wide barrier, so gather latency, tensor-core time and the epilogue overlap across tiles.
#include "common.cuh"
#include "tower_types.cuh"
#include "tower_tc.cuh"
#include "tctile.cuh"
#include <algorithm>

namespace cfm {

constexpr int TC_EPI_WARPS = 4;
constexpr int TC_GROUPS = 2;                                             // producer groups, alternating chunks
constexpr int TC_GROUP_WARPS = 8;
constexpr int TC_PROD_WARPS = TC_GROUPS * TC_GROUP_WARPS;
constexpr int TC_MMA_WARP = TC_EPI_WARPS;
constexpr int TC_PROD_WARP0 = TC_EPI_WARPS + 1;
constexpr int TC_THREADS = (TC_EPI_WARPS + 1 + TC_PROD_WARPS) * 32;      // 672
constexpr int TC_GT = TC_GROUP_WARPS * 32;                               // threads of one producer group
constexpr int TC_PT = TC_PROD_WARPS * 32;                                // all producer threads
constexpr int TCF_M = 128;                                               // forward tile rows (accumulator M)
constexpr int TCB_M = 64;                                                // backward tile rows
constexpr int TC_SMEM_MAX = 227 * 1024;

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// optional event trace (debug): cfm_debug_set_trace(buffer of n_ctas * 128 uint64) -> every CTA records
// globaltimer stamps of its roles; slot layout: [0] start, [1] setup done, [2 + i] producer chunk i arrived (i < 40),
// [48 + t] MMA tile t committed (t < 16), [64 + t] epilogue tile t done (t < 16), [127] end
__device__ unsigned long long* g_trace = nullptr;
__device__ int g_trace_code = 0;                 // 10 * (0 forward, 1 backward) + stage of the launches to record
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void trace(int slot, int code) {
    if (g_trace && g_trace_code == code) g_trace[(size_t)blockIdx.x * 128 + slot] = gtime();
}
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
// Asynchronous global -> shared copies (LDGSTS).  Register loads all land on one hardware scoreboard in these kernels
// (ptxas gives every LDG.128 of the producer loops barrier 5), so waiting for the oldest of several software-
// pipelined loads waits for the youngest too and the pipeline collapses to one load latency per chunk.  cp.async
// groups complete in order and `wait_group N` waits for exactly the groups it names.  bytes < size zero-fills.
__device__ __forceinline__ void cp_async16(void* dst, const void* src, int bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async8(void* dst, const void* src, int bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst, const void* src, int bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void sts4(uint8_t* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }

// ------------------------------------------------------------------------------------------
// weight images
// ------------------------------------------------------------------------------------------
struct PrepTower {
    const float* W[3];
    int K[3], N[3];
    int KE, n_num;
    float* img;
};
struct PrepArgs { PrepTower t[2]; };

__global__ void __launch_bounds__(256) tc_prep_weights(const __grid_constant__ PrepArgs a) {
    const PrepTower& T = a.t[blockIdx.y];
    if (!T.img) return;
    const WImgLayout L = wimg_layout(T.K, T.N);
    for (int s = 0; s < 3; ++s) {
        const int K = T.K[s], N = T.N[s], npad = tc_npad(N), nch = tc_nch(K), ncolp = tc_nblk(N) * 32;
        const int kcols = max(nch * 32, tcb_nch(K) * tcb_cw(K));
        float* wimg = T.img + L.w[s];
        float* wtimg = T.img + L.wt[s];
        const int total = ncolp * kcols;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
            const int n = i / kcols, kc = i - n * kcols;
            float v = 0.f;
            if (n < N && kc < K) v = __ldg(T.W[s] + (size_t)n * K + (s == 0 ? gcol_stage1(kc, T.KE, T.n_num) : kc));
            float hi, lo;
            split_tf32(v, hi, lo);
            const int j = kc >> 5, c = kc & 31;
            if (n < npad && kc < nch * 32) {
                float* ch = wimg + (size_t)j * tc_w_chunk_floats(N);
                const uint32_t off = sw128_off(n, c, npad) >> 2;
                ch[off] = hi;
                ch[npad * 32 + off] = lo;
            }
            const int cw = tcb_cw(K);
            float* tch = wtimg + (size_t)(kc / cw) * tc_wt_chunk_floats(K, N);
            const uint32_t offt = sw128_off(kc % cw, n, cw) >> 2;
            tch[offt] = hi;
            tch[cw * ncolp + offt] = lo;
        }
    }
}

// ------------------------------------------------------------------------------------------
// producers: one chunk = 32 input columns of a row tile
// ------------------------------------------------------------------------------------------
// a = dropout(relu(bn(h))) for a 4-column quad.  `code` is what the forward saves for the backward: the normalised
// pre-activation x-hat (the raw pre-activation when no BatchNorm precedes) where the unit is active and kept, NaN
// elsewhere - one tensor from which the backward recovers a, the activation derivative and x-hat (act_decode).
__device__ __forceinline__ float4 act_quad(const float4& h, const float* sm_bn, int Kp, int c0, int bn_mode, const DropCtx& drop,
                                           long long row, float4* code) {
    float t[4] = {h.x, h.y, h.z, h.w}, u[4] = {h.x, h.y, h.z, h.w};
    if (bn_mode) {
        const float4 m = *reinterpret_cast<const float4*>(sm_bn + c0), is = *reinterpret_cast<const float4*>(sm_bn + Kp + c0);
        const float4 ga = *reinterpret_cast<const float4*>(sm_bn + 2 * Kp + c0), be = *reinterpret_cast<const float4*>(sm_bn + 3 * Kp + c0);
        u[0] = (t[0] - m.x) * is.x; u[1] = (t[1] - m.y) * is.y; u[2] = (t[2] - m.z) * is.z; u[3] = (t[3] - m.w) * is.w;
        t[0] = u[0] * ga.x + be.x; t[1] = u[1] * ga.y + be.y; t[2] = u[2] * ga.z + be.z; t[3] = u[3] * ga.w + be.w;
    }
    bool keep[4] = {true, true, true, true};
    if (drop.active) {
        const Philox4 w = drop_words(drop, row, c0 >> 2);
        keep[0] = w.x >= drop.thresh; keep[1] = w.y >= drop.thresh; keep[2] = w.z >= drop.thresh; keep[3] = w.w >= drop.thresh;
    }
    const float nanv = __int_as_float(0x7fc00000);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const bool on = t[e] > 0.f && keep[e];
        t[e] = on ? t[e] * drop.inv_keep : 0.f;
        u[e] = on ? u[e] : nanv;
    }
    if (code) *code = make_float4(u[0], u[1], u[2], u[3]);
    return make_float4(t[0], t[1], t[2], t[3]);
}
// backward side of act_quad: from the saved code, a (for the weight gradient), d a / d pre-activation and x-hat
__device__ __forceinline__ void act_decode(const float4& code, const float* sm_bn, int Kp, int c0, int bn_mode, float inv_keep,
                                           float4& a, float4& dact, float4& xhat) {
    const float u[4] = {code.x, code.y, code.z, code.w};
    float t[4] = {u[0], u[1], u[2], u[3]};
    if (bn_mode) {
        const float4 ga = *reinterpret_cast<const float4*>(sm_bn + 2 * Kp + c0), be = *reinterpret_cast<const float4*>(sm_bn + 3 * Kp + c0);
        t[0] = u[0] * ga.x + be.x; t[1] = u[1] * ga.y + be.y; t[2] = u[2] * ga.z + be.z; t[3] = u[3] * ga.w + be.w;
    }
    float av[4], dv[4], xv[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const bool on = u[e] == u[e];                 // not NaN
        av[e] = on ? t[e] * inv_keep : 0.f;
        dv[e] = on ? inv_keep : 0.f;
        xv[e] = on ? u[e] : 0.f;
    }
    a = make_float4(av[0], av[1], av[2], av[3]);
    dact = make_float4(dv[0], dv[1], dv[2], dv[3]);
    xhat = make_float4(xv[0], xv[1], xv[2], xv[3]);
}

__device__ __forceinline__ void stage_bn_params_tc(const ActSrc& a, int K, float* sm_bn, int tid, int nthreads) {
    const int Kp = (K + 3) & ~3;
    if (a.bn_mode == 0) return;
    for (int c = tid; c < Kp; c += nthreads) {
        const bool ok = c < K;
        const float s = ok ? a.var_or_istd[c] : 1.f;
        sm_bn[c] = ok ? a.mean[c] : 0.f;
        sm_bn[Kp + c] = a.bn_mode == 2 ? rsqrtf(s + BN_EPS) : s;
        sm_bn[2 * Kp + c] = ok ? a.gamma[c] : 0.f;
        sm_bn[3 * Kp + c] = ok ? a.beta[c] : 0.f;
    }
}

// Position of a chunk in a producer group's stream.  The CTA's chunks, numbered tile-major (c = it * nch + j), go to
// the groups in turn: group g takes c = g, g + TC_GROUPS, ...
struct ChunkPos {
    int it, j;
    __device__ __forceinline__ void start(int grp, int nch) { it = grp / nch; j = grp - it * nch; }
    template <int GROUPS>
    __device__ __forceinline__ void next(int nch) {
        j += GROUPS;
        while (j >= nch) { j -= nch; ++it; }
    }
};

// Software pipeline of a producer thread over its group's chunks: the data loads of the next PD chunks are in flight
// in registers (NX 16-byte quads per chunk), the index loads they depend on two chunks further ahead.  The functors
// must treat positions past the end (it >= number of tiles) as no-ops that still define their outputs.
template <int GROUPS, int PD, int NX, int NI, class IssueIdx, class IssueData, class Consume>
__device__ __forceinline__ void chunk_pipeline(int n_tiles, int nch, int grp, IssueIdx issue_idx, IssueData issue_data,
                                               Consume consume) {
    static_assert(PD % 2 == 0, "the two index sets alternate with the unrolled slot");
    ChunkPos pc, pd, pi;
    pc.start(grp, nch); pd = pc; pi = pc;
    float4 buf[PD][NX];
    long long ix[2][NI];
    {
        long long ixp[PD][NI];
#pragma unroll
        for (int u = 0; u < PD; ++u) { issue_idx(pi, ixp[u]); pi.template next<GROUPS>(nch); }
#pragma unroll
        for (int u = 0; u < PD; ++u) { issue_data(pd, ixp[u], buf[u]); pd.template next<GROUPS>(nch); }
    }
    issue_idx(pi, ix[0]); pi.template next<GROUPS>(nch);
    issue_idx(pi, ix[1]); pi.template next<GROUPS>(nch);
    while (pc.it < n_tiles) {
#pragma unroll
        for (int u = 0; u < PD; ++u) {
            if (pc.it < n_tiles) {
                consume(pc, buf[u]); pc.template next<GROUPS>(nch);
                issue_data(pd, ix[u & 1], buf[u]); pd.template next<GROUPS>(nch);
                issue_idx(pi, ix[u & 1]); pi.template next<GROUPS>(nch);
            }
        }
    }
}

// loads of one stage-1 chunk: 4-column quad q of rows sub + 32 * i (tile order: embedding columns, then numerics)
template <int NX>
__device__ __forceinline__ void gather_issue_idx(const GatherSrc& g, long long B, long long row0, int j, int q, int sub,
                                                 long long (&ix)[NX]) {
    const int c0 = 32 * j + 4 * q;
    if (c0 < g.n_tab * g.E) {
        const int t = c0 / g.E;
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            const long long r = row0 + sub + 32 * i;
            if (r < B) ix[i] = __ldg(g.x_cat + r * g.n_tab + t);
        }
    }
}
template <int NX>
__device__ __forceinline__ void gather_issue_data(const GatherSrc& g, long long B, long long row0, int j, int q, int sub,
                                                  const long long (&ix)[NX], int* err, float4* x) {
    const int c0 = 32 * j + 4 * q;
    const int KE = g.n_tab * g.E;
    if (c0 < KE) {                           // E % 4 == 0: a quad never straddles two tables
        const int t = c0 / g.E, e = c0 - t * g.E;
        const float* tb = g.tab[t];
        const long long trows = g.tab_rows[t];
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            const long long r = row0 + sub + 32 * i;
            if (r < B) {
                long long v = ix[i];
                if (v < 0 || v >= trows) { if (err) atomicOr(err, CFM_FLAG_INDEX_OOB); v = 0; }
                x[i] = ldg4(tb + (size_t)v * g.E + e);
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            const long long r = row0 + sub + 32 * i;
            if (r < B) {
                float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int cn = c0 + e - KE;
                    if (cn < g.n_num) v[e] = __ldg(g.x_num + (size_t)r * g.n_num + cn);
                }
                x[i] = make_float4(v[0], v[1], v[2], v[3]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// forward stage
// ------------------------------------------------------------------------------------------
constexpr int TCF_PD = 2;                                // chunks of input a producer group keeps in flight
constexpr int TCF_NIS = TCF_PD + 1;                      // index slots (stage 1): one being read, TCF_PD in flight
constexpr int TCF_RAW_SLOT = (TCF_M / 32) * TC_GT * 16;  // one chunk of one group: NX quads per thread, 16 KB
constexpr int TCF_IDX_SLOT = (TCF_M / 32) * TC_GT * 8;   // its embedding indices, 8 KB
struct TcFwdSmem {
    int w, raw, idx, ring, bn, bias, piv, wst, bars, tmem, total, nst, stage, wstream;
};
// Shared-memory plan of the forward stage.  `stage1`: the input is gathered (index staging needed).  The weight image
// stays resident when at least two ring stages fit beside it; otherwise (wide stage-1 inputs: 7 chunks x 16 KB for the
// firm tower of config 4) every ring stage carries its own weight chunk, re-streamed from L2 per tile.
__host__ __device__ inline TcFwdSmem tcf_smem(int K, int N, bool stage1) {
    TcFwdSmem s;
    const int nch = tc_nch(K), Kp = (K + 3) & ~3;
    const int chunk_bytes = tc_w_chunk_floats(N) * 4, img = 2 * TCF_M * 128;
    int f = 0;                                          // fixed tail (relative)
    const int bn = f; f += 4 * Kp * 4;
    const int bias = f; f += 64 * 4;
    const int piv = f; f += 4 * 64 * 4;
    const int wst = f; f += 4 * (2 * 64 + 4) * 4;
    f = (f + 15) & ~15;
    const int bars = f; f += 32 * 8;
    const int tmem = f; f += 16;
    const int raw_bytes = TC_GROUPS * TCF_PD * TCF_RAW_SLOT, idx_bytes = stage1 ? TC_GROUPS * TCF_NIS * TCF_IDX_SLOT : 0;
    const int avail = TC_SMEM_MAX - 1024 - raw_bytes - idx_bytes - f;
    int nst = (avail - nch * chunk_bytes) / img;
    s.wstream = nst < 2;
    s.stage = s.wstream ? img + chunk_bytes : img;
    if (s.wstream) nst = avail / s.stage;
    nst = nst > 4 ? 4 : nst;
    s.nst = nst;
    int o = 0;
    s.w = o; o += s.wstream ? 0 : nch * chunk_bytes;
    s.raw = o; o += raw_bytes;
    s.idx = o; o += idx_bytes;
    s.ring = o; o += (nst > 0 ? nst : 0) * s.stage;
    s.bn = o + bn; s.bias = o + bias; s.piv = o + piv; s.wst = o + wst; s.bars = o + bars; s.tmem = o + tmem;
    s.total = o + f + 1024;
    return s;
}

template <int NC, bool STAGE1>
__global__ void __launch_bounds__(TC_THREADS, 1) tower_fwd_tc(const __grid_constant__ FwdArgs args) {
    // the towers share the SMs side by side: CTAs [0, cta_split) belong to tower 0, the others to tower 1
    const int tower = (int)blockIdx.x >= args.cta_split;
    const int cta = tower ? (int)blockIdx.x - args.cta_split : (int)blockIdx.x;
    const int nctas = tower ? (int)gridDim.x - args.cta_split : args.cta_split;
    const FwdStage& S = args.st[tower];
    extern __shared__ uint8_t smem_raw[];
    uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic on the __shared__ array: accesses compile to LDS/STS (a uintptr_t round trip makes them generic LD/ST)
    const int K = S.in.K, N = S.N, npad = tc_npad(N), nch = tc_nch(K), Kp = (K + 3) & ~3;
    const TcFwdSmem L = tcf_smem(K, N, STAGE1);
    const int nst = L.nst;
    const bool wstream = L.wstream != 0;
    float* sm_bn = reinterpret_cast<float*>(sm + L.bn);
    float* sm_bias = reinterpret_cast<float*>(sm + L.bias);
    float* sm_piv = reinterpret_cast<float*>(sm + L.piv);
    float* sm_wst = reinterpret_cast<float*>(sm + L.wst);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L.bars);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + L.tmem);
    uint64_t *w_full = bars, *full = bars + 1, *empty = bars + 5, *d_full = bars + 9, *d_empty = bars + 11;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long B = args.B;
    const long long ntiles = (B + TCF_M - 1) / TCF_M;
    const int my_tiles = (long long)cta < ntiles ? (int)((ntiles - cta + nctas - 1) / nctas) : 0;
    const bool exact = args.exact != 0;
    const bool stats = S.stat_part != nullptr;
    const int chunk_bytes = tc_w_chunk_floats(N) * 4;
    constexpr int IMG_BYTES = TCF_M * 128;
    const int STAGE_BYTES = L.stage;                       // (hi, lo) input images [+ the chunk's weight image]
    const int tcode = S.in.stage;

    if (tid == 0) {
        trace(0, tcode);
        mbar_init(w_full, 1);
        // a stage is full when the group's 256 threads have written the images and its weight chunk is in place
        for (int s = 0; s < 4; ++s) { mbar_init(full + s, TC_GT + 1); mbar_init(empty + s, 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(d_full + b, 1); mbar_init(d_empty + b, TC_EPI_WARPS * 32); }
        fence_barrier_init();
    }
    constexpr uint32_t TMEM_COLS = 2 * NC;
    if (warp == TC_MMA_WARP) tmem_alloc(tmem_slot, TMEM_COLS);
    if (!STAGE1) stage_bn_params_tc(S.in.a, K, sm_bn, tid, TC_THREADS);
    for (int c = tid; c < 64; c += TC_THREADS) sm_bias[c] = c < N ? S.bias[c] : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (tid == 0) trace(1, tcode);

    if (warp == TC_MMA_WARP) {
        // ===================== MMA issuer =====================
        // The whole warp walks the loops (warp-uniform control flow keeps descriptors in uniform registers); one
        // elected lane issues the tcgen05 instructions.
        if (my_tiles > 0) {
            if (!wstream) {
                if (elect_one()) {
                    mbar_expect_tx(w_full, (uint32_t)(nch * chunk_bytes));
                    for (int j = 0; j < nch; ++j)
                        bulk_g2s(sm + L.w + j * chunk_bytes, S.wimg + (size_t)j * tc_w_chunk_floats(N), (uint32_t)chunk_bytes, w_full);
                }
                __syncwarp();
                mbar_wait(w_full, 0);
            }
            const uint32_t idesc = make_idesc_tf32(TCF_M, npad, false, false);
            const uint32_t w_s = smem_u32(sm + L.w), ring_s = smem_u32(sm + L.ring);
            const int ks_last = (K - 32 * (nch - 1) + 7) >> 3;
            const int p0 = exact ? 0 : 2;
            uint32_t cnt = 0;
            for (int it = 0; it < my_tiles; ++it) {
                const int buf = it & 1;
                mbar_wait(d_empty + buf, ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + buf * NC;
                for (int j = 0; j < nch; ++j, ++cnt) {
                    const int s = cnt % nst;
                    mbar_wait(full + s, (cnt / nst) & 1);
                    tc_fence_after();
                    const uint32_t a_hi = ring_s + s * STAGE_BYTES, a_lo = a_hi + IMG_BYTES;
                    const uint32_t b_hi = wstream ? a_hi + 2 * IMG_BYTES : w_s + j * chunk_bytes, b_lo = b_hi + npad * 128;
                    const int ksn = j == nch - 1 ? ks_last : 4;
                    if (elect_one()) {
                        for (int p = p0; p < 3; ++p) {
                            const uint32_t ai = p == 0 ? a_lo : a_hi, bi = p == 1 ? b_lo : b_hi;
                            for (int ks = 0; ks < ksn; ++ks)
                                umma_tf32(d, desc_kmajor_sw128(ai + ks * 32), desc_kmajor_sw128(bi + ks * 32), idesc,
                                          !(j == 0 && p == p0 && ks == 0));
                        }
                        umma_commit(empty + s);
                        if (j == nch - 1) umma_commit(d_full + buf);
                    }
                    __syncwarp();
                }
                if (lane == 0 && it < 16) trace(48 + it, tcode);
            }
        }
    } else if (warp < TC_EPI_WARPS) {
        // ===================== epilogue: thread <-> row =====================
        const int q = warp;
        constexpr int NH = NC / 32;                        // 32-column halves; after the reduction lane l owns
        float S1[NH], S2[NH];                              // column 32 * h + l of half h
#pragma unroll
        for (int k = 0; k < NH; ++k) { S1[k] = 0.f; S2[k] = 0.f; }
        float cntw = 0.f;
        bool have_piv = false;
        float* piv = sm_piv + q * 64;
        for (int it = 0; it < my_tiles; ++it) {
            const long long row0 = ((long long)cta + (long long)it * nctas) * TCF_M;
            const int rows_valid = (int)min((long long)TCF_M, B - row0);
            const int buf = it & 1;
            const int r = 32 * q + lane;
            const bool valid = r < rows_valid;
            const int nvw = max(0, min(32, rows_valid - 32 * q));           // valid rows of this warp (warp-uniform)
            mbar_wait(d_full + buf, (it >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int h = 0; h < NH; ++h) {
                float v[32];
                tmem_ld32(tmem_base + buf * NC + 32 * h + ((uint32_t)(32 * q) << 16), v);
                if (h == NH - 1) {                        // accumulator drained: the MMAs of tile it + 2 may start
                    tc_fence_before();
                    mbar_arrive(d_empty + buf);
                }
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float4 b4 = *reinterpret_cast<const float4*>(sm_bias + 32 * h + i);
                    v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
                }
                if (valid) {
                    float* dst = S.hout + (size_t)(row0 + r) * N + 32 * h;
                    if ((N & 3) == 0) {
#pragma unroll
                        for (int i = 0; i < 32; i += 4)
                            if (32 * h + i < N) *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (32 * h + i < N) dst[i] = v[i];
                    }
                }
                if (stats && nvw > 0) {
                    if (!have_piv) {
                        // pivot = column means of this warp's first rows: keeps the shifted sums free of cancellation
                        float t[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) t[i] = valid ? v[i] : 0.f;
                        Halve<32, 16>::run(t, lane);
                        piv[32 * h + lane] = t[0] / (float)nvw;
                        __syncwarp();
                    }
                    float e[32];
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        const float4 p4 = *reinterpret_cast<const float4*>(piv + 32 * h + i);
                        v[i] = valid ? v[i] - p4.x : 0.f; v[i + 1] = valid ? v[i + 1] - p4.y : 0.f;
                        v[i + 2] = valid ? v[i + 2] - p4.z : 0.f; v[i + 3] = valid ? v[i + 3] - p4.w : 0.f;
                    }
#pragma unroll
                    for (int i = 0; i < 32; ++i) e[i] = v[i] * v[i];
                    Halve<32, 16>::run(v, lane);
                    Halve<32, 16>::run(e, lane);
                    S1[h] += v[0];
                    S2[h] += e[0];
                }
            }
            if (stats && nvw > 0) { have_piv = true; cntw += (float)nvw; }
            if (tid == 0 && it < 16) trace(64 + it, tcode);
        }
        if (stats) {
            // per-warp (count, mean, M2) -> Chan merge of the four warps in warp order -> the CTA's partial
            float* ws = sm_wst + q * (2 * 64 + 4);
#pragma unroll
            for (int k = 0; k < NH; ++k) {
                const int c = 32 * k + lane;
                const float pv = have_piv ? piv[c] : 0.f;
                ws[c] = cntw > 0.f ? pv + S1[k] / cntw : 0.f;
                ws[64 + c] = cntw > 0.f ? fmaxf(S2[k] - S1[k] * S1[k] / cntw, 0.f) : 0.f;
            }
            if (lane == 0) ws[128] = cntw;
            named_bar_sync(1, TC_EPI_WARPS * 32);
            float* P = S.stat_part + (size_t)cta * (2 * N + 4);
            const int c = tid;
            if (c < N) {
                float n = 0.f, mean = 0.f, m2 = 0.f;
#pragma unroll
                for (int w = 0; w < TC_EPI_WARPS; ++w) {
                    const float* o = sm_wst + w * (2 * 64 + 4);
                    const float nb = o[128];
                    if (nb > 0.f) {
                        const float d = o[c] - mean, nn = n + nb;
                        mean += d * (nb / nn);
                        m2 += o[64 + c] + d * d * (n * nb / nn);
                        n = nn;
                    }
                }
                P[c] = mean;
                P[N + c] = m2;
                if (c == 0) P[2 * N] = n;
            }
        }
    } else {
        // ===================== producers =====================
        // Per group and chunk: (A) request the chunk's embedding indices, (B) TCF_PD chunks later read them and request
        // the data quads, (C) TCF_PD chunks later read the quads back, build the (hi, lo) images, hand the stage over.
        // Everything in flight sits in thread-private shared-memory slots filled by cp.async; every iteration commits
        // exactly two groups (indices, then data), so a constant `wait_group` count names the group each step needs.
        // Order inside an iteration: A, C, B - the data slot (B) refills is the one (C) has just read.
        const int ptid = tid - TC_PROD_WARP0 * 32;
        const int grp = ptid / TC_GT, gtid = ptid - grp * TC_GT;
        const int q = gtid & 7, sub = gtid >> 3;
        const DropCtx drop = resolve_drop(S.in.a.drop);
        uint8_t* ring = sm + L.ring;
        const GatherSrc& g = S.in.g;
        constexpr int NX = TCF_M / 32;
        const int KE = STAGE1 ? g.n_tab * g.E : 0;
        uint8_t* raw = sm + L.raw + grp * TCF_PD * TCF_RAW_SLOT + gtid * 16;        // + (slot * NX + i) * TC_GT * 16
        uint8_t* ixs = sm + L.idx + grp * TCF_NIS * TCF_IDX_SLOT + gtid * 8;        // + (slot * NX + i) * TC_GT * 8
        auto tile_row0 = [&](int it) { return ((long long)cta + (long long)it * nctas) * TCF_M; };
        auto issue_idx = [&](const ChunkPos& p, int slot) {
            if (!STAGE1 || p.it >= my_tiles) return;
            const int c0 = 32 * p.j + 4 * q;
            if (c0 >= KE) return;                        // numeric columns need no index
            const long long row0 = tile_row0(p.it);
            const int t = c0 / g.E;
#pragma unroll
            for (int i = 0; i < NX; ++i) {
                const long long r = row0 + sub + 32 * i;
                cp_async8(ixs + (slot * NX + i) * TC_GT * 8, g.x_cat + (r < B ? r : 0) * g.n_tab + t, r < B ? 8 : 0);
            }
        };
        auto issue_data = [&](const ChunkPos& p, int islot, int dslot) {
            if (p.it >= my_tiles) return;
            const long long row0 = tile_row0(p.it);
            const int c0 = 32 * p.j + 4 * q;
            uint8_t* dst = raw + dslot * NX * TC_GT * 16;
            if (!STAGE1) {
#pragma unroll
                for (int i = 0; i < NX; ++i) {
                    const long long r = row0 + sub + 32 * i;
                    const bool ok = r < B && c0 < K;
                    cp_async16(dst + i * TC_GT * 16, S.in.a.h + (ok ? (size_t)r * K + c0 : 0), ok ? 16 : 0);
                }
            } else if (c0 < KE) {                        // E % 4 == 0: a quad never straddles two tables
                const int t = c0 / g.E, e = c0 - t * g.E;
                const float* tb = g.tab[t];
                const long long trows = g.tab_rows[t];
#pragma unroll
                for (int i = 0; i < NX; ++i) {
                    const long long r = row0 + sub + 32 * i;
                    long long v = *reinterpret_cast<const long long*>(ixs + (islot * NX + i) * TC_GT * 8);
                    if (r < B && (v < 0 || v >= trows)) { if (args.err) atomicOr(args.err, CFM_FLAG_INDEX_OOB); v = 0; }
                    cp_async16(dst + i * TC_GT * 16, tb + (r < B ? (size_t)v * g.E + e : 0), r < B ? 16 : 0);
                }
            } else {                                     // numeric columns (possibly a ragged quad): 4-byte pieces
#pragma unroll
                for (int i = 0; i < NX; ++i) {
                    const long long r = row0 + sub + 32 * i;
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int cn = c0 + e - KE;
                        const bool ok = r < B && cn < g.n_num;
                        cp_async4(dst + i * TC_GT * 16 + 4 * e, g.x_num + (ok ? (size_t)r * g.n_num + cn : 0), ok ? 4 : 0);
                    }
                }
            }
        };
        auto consume = [&](const ChunkPos& p, int dslot) {
            const uint32_t cnt = (uint32_t)(p.it * nch + p.j);
            const int s = cnt % nst;
            mbar_wait(empty + s, ((cnt / nst) & 1) ^ 1);
            const long long row0 = tile_row0(p.it);
            const int c0 = 32 * p.j + 4 * q;
            uint8_t* hi_img = ring + s * STAGE_BYTES;
            if (gtid == 0) {                             // the stage's 257th arrival: its weight chunk (or nothing)
                if (wstream) {
                    mbar_expect_tx(full + s, (uint32_t)chunk_bytes);
                    bulk_g2s(hi_img + 2 * IMG_BYTES, S.wimg + (size_t)p.j * tc_w_chunk_floats(N), (uint32_t)chunk_bytes, full + s);
                } else {
                    mbar_arrive(full + s);
                }
            }
            const uint8_t* src = raw + dslot * NX * TC_GT * 16;
#pragma unroll
            for (int i = 0; i < NX; ++i) {
                const int r = sub + 32 * i;
                float4 a = *reinterpret_cast<const float4*>(src + i * TC_GT * 16);
                if (STAGE1 && S.x_out && row0 + r < B) {
                    // input stash for the backward: 16 KB blocks [64-row tile][64-column chunk], rows of 16 quads with
                    // the quad index XOR-ed by the row (the backward reads it lane <-> row without bank conflicts)
                    const long long gr = row0 + r;
                    const int q16 = (c0 & 63) >> 2, r64 = (int)(gr & 63);
                    float* blk = S.x_out + ((gr >> 6) * tcb_nch(K) + (c0 >> 6)) * 4096;
                    *reinterpret_cast<float4*>(blk + r64 * 64 + ((q16 ^ (r64 & 15)) << 2)) = a;
                }
                if (!STAGE1) {
                    if (row0 + r < B && c0 < K) {
                        float4 code;
                        a = act_quad(a, sm_bn, Kp, c0, S.in.a.bn_mode, drop, row0 + r, &code);
                        if (S.a_out) *reinterpret_cast<float4*>(S.a_out + (size_t)(row0 + r) * K + c0) = code;
                    } else {
                        a = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
                float4 hi, lo;
                split_tf32x4(a, hi, lo);
                const uint32_t off = sw128_chunk(r, 0, q, TCF_M);
                sts4(hi_img + off, hi);
                if (exact) sts4(hi_img + IMG_BYTES + off, lo);
            }
            fence_proxy_async();
            mbar_arrive(full + s);
            if (gtid == 0 && cnt < 40) trace(2 + cnt, tcode);
        };
        ChunkPos pi, pd, pc;
        pi.start(grp, nch); pd = pi; pc = pi;
        int si = 0, sd_i = 0, sd_d = 0, sc_d = 0;        // slot counters: index write / index read / data write / data read
        for (int jj = -2 * TCF_PD;; ++jj) {
            if (jj >= 0 && pc.it >= my_tiles) break;
            issue_idx(pi, si); cp_async_commit();                                        // (A) indices of chunk jj + 2 PD
            pi.template next<TC_GROUPS>(nch); si = si + 1 == TCF_NIS ? 0 : si + 1;
            if (jj >= 0) {
                cp_async_wait<2 * TCF_PD - 1>();         // (C) the data of chunk jj have landed (and every older group)
                consume(pc, sc_d);
                pc.template next<TC_GROUPS>(nch); sc_d = sc_d + 1 == TCF_PD ? 0 : sc_d + 1;
            } else if (jj >= -TCF_PD) {
                cp_async_wait<2 * TCF_PD>();             // start-up: the indices of chunk jj + PD have landed
            }
            if (jj >= -TCF_PD) {                         // (B) data of chunk jj + PD, into the slot (C) just drained
                issue_data(pd, sd_i, sd_d);
                pd.template next<TC_GROUPS>(nch);
                sd_i = sd_i + 1 == TCF_NIS ? 0 : sd_i + 1; sd_d = sd_d + 1 == TCF_PD ? 0 : sd_d + 1;
            }
            cp_async_commit();
        }
        cp_async_wait<0>();
    }
    tc_fence_before();
    __syncthreads();
    if (tid == 0) trace(127, tcode);
    if (warp == TC_MMA_WARP) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------
// backward stage: per 64-row tile  dX = G . W  (chunk by chunk of CW = 64 input columns; M = 64 accumulators double
// buffered in TMEM) and  dW += G^T . [A | 1]  (accumulated in TMEM over ALL tiles of the CTA, written once).
// Every tcgen05.mma costs about 60 cycles to issue whatever its shape, so the chunks are as wide as shared memory
// allows, and the bias gradient rides in the last chunk's product as a ninth row group (a constant row of ones).
// Producer threads own one batch row each (lane <-> row), so the transposed images G^T and A^T are written straight
// from registers, one 128-byte row per store instruction.
// ------------------------------------------------------------------------------------------
struct TcBwdSmem {
    int g, gt, raw, slots, ring, at, wt, msk, stage, bna, bng, red, bars, tmem, total, nst, at_img;
};
__host__ __device__ inline TcBwdSmem tcb_smem(int K, int N, bool stage1, bool a_bn, bool stash = false) {
    TcBwdSmem s;
    const int nbn = tc_nblk(N), Kp = (K + 3) & ~3, Np = (N + 3) & ~3, cw = tcb_cw(K);
    int o = 0;
    s.g = o; o += 2 * TCB_M * 128 * nbn;              // G   [64 rows][N]      hi, lo
    s.gt = o; o += 2 * 64 * 128 * 2;                  // G^T [64 n][64 rows]   hi, lo
    s.raw = o; o += stash ? 64 * 64 * 4 : 0;          // one block of the forward's input stash (bulk copy target)
    s.slots = o; o += stage1 ? 0 : 2 * (cw / 32) * TC_PT * 16;   // stage > 1: two chunks of saved codes in flight (cp.async)
    s.ring = o;
    int st = 0;
    s.at_img = (cw + 8) * 128 * 2;                    // A^T chunk [cw k + 8][64 rows]: row cw = ones (bias gradient)
    s.at = st; st += 2 * s.at_img;
    s.wt = st; st += 2 * cw * 128 * nbn;              // W^T chunk [cw k][N]       hi, lo
    s.msk = st; st += stage1 ? 0 : TCB_M * cw * 4;    // the chunk's activation codes [64 rows][cw k]: the epilogue decodes
    (void)a_bn;                                       // d(act)/d(pre-activation) and x-hat from them (act_quad)
    s.stage = st;
    int f = 0;
    const int bna = f; f += 4 * Kp * 4;
    const int bng = f; f += 5 * Np * 4;
    const int red = f; f += 4 * 2 * 64 * 4;
    f = (f + 15) & ~15;
    const int bars = f; f += 32 * 8;
    const int tmem = f; f += 16;
    int nst = (TC_SMEM_MAX - 1024 - o - f) / st;
    nst = nst > 4 ? 4 : nst;
    s.nst = nst;
    o += (nst > 0 ? nst : 0) * st;
    s.bna = o + bna; s.bng = o + bng; s.red = o + red; s.bars = o + bars; s.tmem = o + tmem;
    s.total = o + f + 1024;
    return s;
}
__host__ __device__ inline uint32_t tcb_tmem_cols(int K) {
    const uint32_t need = tcb_nch(K) * tcb_cw(K) + 8 + 2 * tcb_cw(K);
    uint32_t c = 32;
    while (c < need) c <<= 1;
    return c;
}

template <bool STAGE1, int CW>
__global__ void __launch_bounds__(TC_THREADS, 1) tower_bwd_tc(const __grid_constant__ BwdArgs args) {
    const int tower = (int)blockIdx.x >= args.cta_split;       // see tower_fwd_tc
    const int cta = tower ? (int)blockIdx.x - args.cta_split : (int)blockIdx.x;
    const int nctas = tower ? (int)gridDim.x - args.cta_split : args.cta_split;
    const BwdStage& S = args.st[tower];
    extern __shared__ uint8_t smem_raw[];
    uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic on the __shared__ array: accesses compile to LDS/STS (a uintptr_t round trip makes them generic LD/ST)
    const int K = S.in.K, N = S.N, npad = tc_npad(N), nbn = tc_nblk(N), nch = tcb_nch(K);
    const int Kp = (K + 3) & ~3, Np = (N + 3) & ~3;
    const bool a_bn = S.a_bn != 0, need_dx = S.need_dx != 0;
    const bool stash = STAGE1 && S.x_in != nullptr;
    const TcBwdSmem L = tcb_smem(K, N, STAGE1, a_bn, stash);
    const int nst = L.nst;
    uint8_t *G = sm + L.g, *GT = sm + L.gt, *ring = sm + L.ring;
    const int G_IMG = TCB_M * 128 * nbn;
    constexpr int GT_IMG = 64 * 128 * 2, AT_IMG = (CW + 8) * 128 * 2, ATR = CW + 8;
    const int WT_IMG = CW * 128 * nbn;
    float* sm_bna = reinterpret_cast<float*>(sm + L.bna);
    float* sm_bng = reinterpret_cast<float*>(sm + L.bng);
    float* sm_red = reinterpret_cast<float*>(sm + L.red);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L.bars);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + L.tmem);
    uint64_t *full = bars, *empty = bars + 4, *g_full = bars + 8, *g_empty = bars + 9, *dx_full = bars + 10, *dx_empty = bars + 12,
             *dw_full = bars + 14, *raw_full = bars + 15;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long B = args.B;
    const long long ntiles = (B + TCB_M - 1) / TCB_M;
    const int my_tiles = (long long)cta < ntiles ? (int)((ntiles - cta + nctas - 1) / nctas) : 0;
    const bool exact = args.exact != 0;
    const int tcode = 10 + S.in.stage;
    const uint32_t tmem_cols = tcb_tmem_cols(K);
    const uint32_t col_db = nch * CW, col_dx = nch * CW + 8;

    if (tid == 0) {
        trace(0, tcode);
        for (int s = 0; s < 4; ++s) { mbar_init(full + s, TC_PT + 1); mbar_init(empty + s, 1 + TC_EPI_WARPS); }
        mbar_init(g_full, TC_PT); mbar_init(g_empty, 1);
        for (int b = 0; b < 2; ++b) { mbar_init(dx_full + b, 1); mbar_init(dx_empty + b, TC_EPI_WARPS * 32); }
        mbar_init(dw_full, 1);
        mbar_init(raw_full, 1);
        fence_barrier_init();
    }
    if (warp == TC_MMA_WARP) tmem_alloc(tmem_slot, tmem_cols);
    // written once: G^T rows past N (zero), and in every ring stage the eight extra A^T rows (ones, zeros)
    for (int i = tid; i < 2 * GT_IMG / 16; i += TC_THREADS) reinterpret_cast<uint4*>(GT)[i] = make_uint4(0, 0, 0, 0);
    for (int s = 0; s < nst; ++s) {
        uint8_t* at = ring + s * L.stage + L.at;
        for (int i = tid; i < 2 * 8 * 64; i += TC_THREADS) {
            const int img = i >> 9, rr = (i >> 6) & 7, c = i & 63;
            *reinterpret_cast<float*>(at + img * AT_IMG + sw128_off(CW + rr, c, ATR)) = (img == 0 && rr == 0) ? 1.f : 0.f;
        }
    }
    if (!STAGE1) stage_bn_params_tc(S.in.a, K, sm_bna, tid, TC_THREADS);
    if (S.g_mode) {
        for (int c = tid; c < Np; c += TC_THREADS) {
            const bool ok = c < N;
            const float sv = ok ? S.g_var_or_istd[c] : 1.f;
            const float istd = S.g_mode == 2 ? rsqrtf(sv + BN_EPS) : sv;
            sm_bng[c] = ok ? S.g_mean[c] : 0.f;
            sm_bng[Np + c] = istd;
            sm_bng[2 * Np + c] = ok ? S.g_gamma[c] * istd : 0.f;
            sm_bng[3 * Np + c] = (ok && S.g_mode == 1) ? S.g_c1[c] : 0.f;
            sm_bng[4 * Np + c] = (ok && S.g_mode == 1) ? S.g_c2[c] : 0.f;
        }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (tid == 0) trace(1, tcode);

    if (warp == TC_MMA_WARP) {
        // ===================== MMA issuer (whole warp in the loops, one elected lane issues) =====================
        if (my_tiles > 0) {
            const uint32_t idesc_x = make_idesc_tf32(64, CW, false, false), idesc_w = idesc_x, idesc_wb = make_idesc_tf32(64, CW + 8, false, false);
            const uint32_t g_s = smem_u32(G), gt_s = smem_u32(GT), ring_s = smem_u32(ring);
            const int p0 = exact ? 0 : 2;
            const int kn = npad >> 3;                        // K steps of dX (over the layer's outputs)
            uint32_t cnt = 0;
            for (int it = 0; it < my_tiles; ++it) {
                mbar_wait(g_full, it & 1);
                tc_fence_after();
                for (int j = 0; j < nch; ++j, ++cnt) {
                    const int s = cnt % nst;
                    mbar_wait(full + s, (cnt / nst) & 1);
                    tc_fence_after();
                    const uint32_t st = ring_s + s * L.stage;
                    const int b = cnt & 1;
                    if (need_dx) {
                        mbar_wait(dx_empty + b, ((cnt >> 1) & 1) ^ 1);
                        tc_fence_after();
                    }
                    if (elect_one()) {
                        if (need_dx) {
                            const uint32_t d = tmem_base + col_dx + CW * b;
                            for (int p = p0; p < 3; ++p) {
                                const uint32_t ai = g_s + (p == 0 ? G_IMG : 0), bi = st + L.wt + (p == 1 ? WT_IMG : 0);
                                for (int ks = 0; ks < kn; ++ks)
                                    umma_tf32(d, tile_desc_k(ai, TCB_M, ks), tile_desc_k(bi, CW, ks), idesc_x, !(p == p0 && ks == 0));
                            }
                            umma_commit(dx_full + b);
                        }
                        const uint32_t d = tmem_base + CW * j;
                        const uint32_t idw = j == nch - 1 ? idesc_wb : idesc_w;     // last chunk: + the row of ones
                        for (int p = p0; p < 3; ++p) {
                            const uint32_t ai = gt_s + (p == 0 ? GT_IMG : 0), bi = st + L.at + (p == 1 ? AT_IMG : 0);
                            for (int ks = 0; ks < 8; ++ks)
                                umma_tf32(d, tile_desc_k(ai, 64, ks), tile_desc_k(bi, ATR, ks), idw, !(it == 0 && p == p0 && ks == 0));
                        }
                        umma_commit(empty + s);
                        if (j == nch - 1) {
                            umma_commit(g_empty);
                            if (it == my_tiles - 1) umma_commit(dw_full);
                        }
                    }
                    __syncwarp();
                }
                if (lane == 0 && it < 16) trace(48 + it, tcode);
            }
        }
    } else if (warp < TC_EPI_WARPS) {
        // ===================== epilogue: M = 64 accumulators, warp q holds rows 16q .. 16q+15 in its lanes 0-15 ======
        const int q = warp;
        const bool act = lane < 16;
        const uint32_t lane_addr = (uint32_t)(32 * q) << 16;
        const int KE = S.in.g.n_tab * S.in.g.E, n_num = S.in.g.n_num;
        float s1[2][2] = {{0.f, 0.f}, {0.f, 0.f}}, s2[2][2] = {{0.f, 0.f}, {0.f, 0.f}};   // [32-column half][owned column] (stage > 1: K <= 64)
        const float inv_keep_e = (!STAGE1 && S.in.a.drop.active) ? S.in.a.drop.inv_keep : 1.f;
        uint32_t cnt = 0;
        for (int it = 0; it < my_tiles; ++it) {
            const long long row0 = ((long long)cta + (long long)it * nctas) * TCB_M;
            const long long row = row0 + 16 * q + lane;
            const bool valid = act && row < B;
            const int rl = 16 * q + (lane & 15);
            for (int j = 0; j < nch; ++j, ++cnt) {
                const int s = cnt % nst;
                const uint8_t* st = ring + s * L.stage;
                if (need_dx) {
                    const int b = cnt & 1;
                    mbar_wait(dx_full + b, (cnt >> 1) & 1);
                    tc_fence_after();
#pragma unroll
                    for (int hf = 0; hf < CW / 32; ++hf) {
                        float v[32];
                        tmem_ld32(tmem_base + col_dx + CW * b + 32 * hf + lane_addr, v);
                        if (hf == CW / 32 - 1) {
                            tc_fence_before();
                            mbar_arrive(dx_empty + b);
                        }
                        const int cb = CW * j + 32 * hf;          // first input column of this half
                        if (STAGE1) {
                            if (valid) {
#pragma unroll
                                for (int i = 0; i < 32; i += 4) {
                                    const int c = cb + i;
                                    if (c + 3 < KE) {
                                        if (S.dx_emb) *reinterpret_cast<float4*>(S.dx_emb + (size_t)row * KE + c) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                                    } else if (c < K) {
#pragma unroll
                                        for (int e = 0; e < 4; ++e) {
                                            const int ce = c + e;
                                            if (ce < KE) { if (S.dx_emb) S.dx_emb[(size_t)row * KE + ce] = v[i + e]; }
                                            else if (ce < K && S.dx_num) S.dx_num[(size_t)row * n_num + (ce - KE)] = v[i + e];
                                        }
                                    }
                                }
                            }
                        } else {
                            // dy = dX * d(act)/d(pre-activation); the producers left that factor (and x-hat) in the ring stage
                            float xs[32];
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                // saved code: x-hat where the unit was active and kept, NaN elsewhere
                                const float4 c4 = *reinterpret_cast<const float4*>(st + L.msk + sw128_chunk(rl, hf, i, TCB_M));
                                const float cc[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const bool on = cc[e] == cc[e];
                                    v[4 * i + e] *= on ? inv_keep_e : 0.f;
                                    xs[4 * i + e] = on ? cc[e] : 0.f;
                                }
                            }
                            if (valid) {
                                float* dst = S.dy_out + (size_t)row * K + cb;
#pragma unroll
                                for (int i = 0; i < 32; i += 4)
                                    if (cb + i < K) *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                            }
                            if (a_bn) {
#pragma unroll
                                for (int i = 0; i < 32; ++i) {
                                    v[i] = valid ? v[i] : 0.f;
                                    xs[i] = v[i] * xs[i];
                                }
                                Halve<32, 8>::run(v, lane);
                                Halve<32, 8>::run(xs, lane);
                                const int hh = (cb >> 5) & 1;       // K <= 64: at most two halves in all
                                if (hh == 0) { s1[0][0] += v[0]; s1[0][1] += v[1]; s2[0][0] += xs[0]; s2[0][1] += xs[1]; }
                                else { s1[1][0] += v[0]; s1[1][1] += v[1]; s2[1][0] += xs[0]; s2[1][1] += xs[1]; }
                            }
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(empty + s);       // this warp no longer reads the stage's mask / x-hat tiles
            }
            if (tid == 0 && it < 16) trace(64 + it, tcode);
        }
        // ---- BatchNorm-backward sums of this CTA: four warps (16 rows each) added in warp order ----
        if (a_bn && need_dx) {
            if (act) {
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        sm_red[(q * 2 + 0) * 64 + 32 * j + 2 * lane + k] = s1[j][k];
                        sm_red[(q * 2 + 1) * 64 + 32 * j + 2 * lane + k] = s2[j][k];
                    }
            }
            named_bar_sync(1, TC_EPI_WARPS * 32);
            float* P = S.sum_part + (size_t)cta * 2 * K;
            if (tid < 2 * K) {
                const int which = tid >= K, c = tid - which * K;
                float t = 0.f;
#pragma unroll
                for (int w = 0; w < TC_EPI_WARPS; ++w) t += sm_red[(w * 2 + which) * 64 + c];
                P[which * K + c] = t;
            }
        }
        // ---- weight-gradient partial of this CTA: TMEM [64 n][K (+ bias column)] -> global, layer column order ----
        {
            float* P = S.dW_part + (size_t)cta * N * (K + 1);
            const int n = 16 * q + lane;
            if (my_tiles > 0) {
                mbar_wait(dw_full, 0);
                tc_fence_after();
            }
            const int ngrp = (int)(col_db + 8 + 31) / 32;    // 32-column groups covering the weight and bias columns
            for (int gidx = 0; gidx < ngrp; ++gidx) {
                float v[32];
                if (my_tiles > 0) {
                    if (32 * gidx + 32 <= (int)tmem_cols) tmem_ld32(tmem_base + 32 * gidx + lane_addr, v);
                    else {
                        float w8[16];
                        tmem_ld16(tmem_base + 32 * gidx + lane_addr, w8);
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = i < 16 ? w8[i] : 0.f;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = 0.f;
                }
                if (act && n < N) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int c = 32 * gidx + i;
                        if (c < K) P[(size_t)n * (K + 1) + (STAGE1 ? gcol_stage1(c, KE, n_num) : c)] = v[i];
                        else if (c == (int)col_db) P[(size_t)n * (K + 1) + K] = v[i];
                    }
                }
            }
            tc_fence_before();
        }
    } else {
        // ===================== producers: thread <-> batch row =====================
        // all sixteen warps work on one chunk at a time: eight threads per batch row
        const int gtid = tid - TC_PROD_WARP0 * 32;
        constexpr int QS = TC_PT / 64;                        // quad stride of a thread (8)
        const int r = gtid & 63, qb = gtid >> 6;              // row of the tile; quads qb, qb + QS, ...
        const GatherSrc& g = S.in.g;
        const int KE = g.n_tab * g.E;
        const float inv_keep = (!STAGE1 && S.in.a.drop.active) ? S.in.a.drop.inv_keep : 1.f;
        constexpr int NXR = CW / (4 * QS);                    // quads per thread and chunk
        constexpr int NX = NXR;
        auto tile_row0 = [&](int it) { return ((long long)cta + (long long)it * nctas) * TCB_M; };

        // G and G^T of tile `it`, built by the group that owns the tile's first chunk.  g = incoming gradient, through
        // the BatchNorm backward of this layer's output when there is one.
        const int qpr = npad >> 2;                            // quads per G row (<= 16)
        // The incoming-gradient quads of a tile are requested one tile ahead (right after the previous G was staged), so
        // that build_g finds them in registers instead of waiting an HBM round trip per tile.
        constexpr int GQ = 16 / QS;                           // at most 16 quads per G row
        float4 gq[GQ], hq[GQ];
        auto issue_g = [&](int it) {
            const long long row = tile_row0(it) + r;
#pragma unroll
            for (int u = 0; u < GQ; ++u) {
                const int c0 = 4 * (qb + QS * u);
                gq[u] = make_float4(0.f, 0.f, 0.f, 0.f); hq[u] = gq[u];
                if (it < my_tiles && row < B && c0 < N) {
                    gq[u] = ldg4(S.gin + (size_t)row * N + c0);
                    if (S.g_mode == 1) hq[u] = ldg4(S.hs + (size_t)row * N + c0);
                }
            }
        };
        auto build_g = [&](int it) {
            const long long row = tile_row0(it) + r;
#pragma unroll
            for (int u = 0; u < GQ; ++u) {
                const int c0 = 4 * (qb + QS * u);
                if (row < B && c0 < N) {
                    float gv[4] = {gq[u].x, gq[u].y, gq[u].z, gq[u].w};
                    if (S.g_mode == 1) {
                        const float hv[4] = {hq[u].x, hq[u].y, hq[u].z, hq[u].w};
                        // five parameter rows of the quad's columns (N % 4 == 0: 16-byte loads)
                        const float4 p0 = *reinterpret_cast<const float4*>(sm_bng + c0), p1 = *reinterpret_cast<const float4*>(sm_bng + Np + c0);
                        const float4 p2 = *reinterpret_cast<const float4*>(sm_bng + 2 * Np + c0), p3 = *reinterpret_cast<const float4*>(sm_bng + 3 * Np + c0);
                        const float4 p4 = *reinterpret_cast<const float4*>(sm_bng + 4 * Np + c0);
                        const float mu[4] = {p0.x, p0.y, p0.z, p0.w}, is[4] = {p1.x, p1.y, p1.z, p1.w}, gi[4] = {p2.x, p2.y, p2.z, p2.w};
                        const float k1[4] = {p3.x, p3.y, p3.z, p3.w}, k2[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float xh = (hv[e] - mu[e]) * is[e];
                            gv[e] = gi[e] * (gv[e] - k1[e] - xh * k2[e]);
                        }
                    } else if (S.g_mode == 2) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) gv[e] *= sm_bng[2 * Np + c0 + e];
                    }
                    gq[u] = make_float4(gv[0], gv[1], gv[2], gv[3]);
                }
            }
            if (gtid == 0 && it < 6) trace(80 + 4 * it, tcode);
            if (it > 0) mbar_wait(g_empty, (it - 1) & 1);      // the MMAs of the previous tile have read G / G^T
            if (gtid == 0 && it < 6) trace(81 + 4 * it, tcode);
#pragma unroll
            for (int u = 0; u < GQ; ++u) {
                const int qq = qb + QS * u;
                if (qq < qpr) {
                    float4 hi, lo;
                    split_tf32x4(gq[u], hi, lo);
                    const uint32_t off = sw128_chunk(r, qq >> 3, qq & 7, TCB_M);
                    sts4(G + off, hi);
                    if (exact) sts4(G + G_IMG + off, lo);
                    const float h4[4] = {hi.x, hi.y, hi.z, hi.w}, l4[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const uint32_t o2 = sw128_off(4 * qq + e, r, 64);
                        *reinterpret_cast<float*>(GT + o2) = h4[e];
                        if (exact) *reinterpret_cast<float*>(GT + GT_IMG + o2) = l4[e];
                    }
                }
            }
            fence_proxy_async();
            mbar_arrive(g_full);
            if (gtid == 0 && it < 6) trace(83 + 4 * it, tcode);
            issue_g(it + 1);
        };
        issue_g(0);

        uint8_t* raw = sm + L.raw;
        auto stash_block = [&](const ChunkPos& p) {
            return S.x_in + ((tile_row0(p.it) >> 6) * nch + p.j) * 4096;
        };
        if (stash && gtid == 0 && my_tiles > 0) {       // block of the first chunk
            mbar_expect_tx(raw_full, 16384u);
            bulk_g2s(raw, S.x_in + (tile_row0(0) >> 6) * nch * 4096, 16384u, raw_full);
        }
        auto issue_idx = [&](const ChunkPos& p, long long (&ix)[NXR]) {
#pragma unroll
            for (int i = 0; i < NXR; ++i) ix[i] = 0;
            if (STAGE1 && !stash && p.it < my_tiles) {
                const long long row = tile_row0(p.it) + r;
                if (row < B) {
#pragma unroll
                    for (int i = 0; i < NXR; ++i) {
                        const int c0 = CW * p.j + 4 * (qb + QS * i);
                        if (c0 < KE) ix[i] = __ldg(g.x_cat + row * g.n_tab + c0 / g.E);
                    }
                }
            }
        };
        auto issue_data = [&](const ChunkPos& p, const long long (&ix)[NXR], float4 (&x)[NX]) {
#pragma unroll
            for (int i = 0; i < NX; ++i) x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.it >= my_tiles || stash) return;
            const long long row = tile_row0(p.it) + r;
            if (row >= B) return;
#pragma unroll
            for (int i = 0; i < NXR; ++i) {
                const int c0 = CW * p.j + 4 * (qb + QS * i);
                if (STAGE1) {
                    if (c0 < KE) {                       // E % 4 == 0: a quad never straddles two tables
                        const int t = c0 / g.E, e = c0 - t * g.E;
                        long long v = ix[i];
                        if (v < 0 || v >= g.tab_rows[t]) v = 0;      // reported by the forward
                        x[i] = ldg4(g.tab[t] + (size_t)v * g.E + e);
                    } else if (c0 < K) {
                        float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (c0 + e < K) v[e] = __ldg(g.x_num + (size_t)row * g.n_num + (c0 + e - KE));
                        x[i] = make_float4(v[0], v[1], v[2], v[3]);
                    }
                } else if (c0 < K) {
                    x[i] = ldg4(S.in.a.a_post + (size_t)row * K + c0);       // the forward's saved code (act_quad)
                }
            }
        };
        auto consume = [&](const ChunkPos& p, float4 (&x)[NX]) {
            const uint32_t cnt = (uint32_t)(p.it * nch + p.j);
            if (STAGE1 && stash) {
                // this chunk's block of the forward's input stash has been travelling since the previous chunk
                mbar_wait(raw_full, cnt & 1);
                const bool rv = tile_row0(p.it) + r < B;
#pragma unroll
                for (int i = 0; i < NXR; ++i) {
                    const int q = qb + QS * i;
                    x[i] = (rv && CW * p.j + 4 * q < K) ? *reinterpret_cast<const float4*>(raw + r * 256 + ((q ^ (r & 15)) << 4))
                                                        : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                fence_proxy_async();
                named_bar_sync(2, TC_PT);                      // every thread holds its quads: the buffer is free again
                ChunkPos nx = p;
                nx.template next<1>(nch);
                if (gtid == 0 && nx.it < my_tiles) {
                    mbar_expect_tx(raw_full, 16384u);
                    bulk_g2s(raw, stash_block(nx), 16384u, raw_full);
                }
            }
            const int s = cnt % nst;
            mbar_wait(empty + s, ((cnt / nst) & 1) ^ 1);
            uint8_t* st = ring + s * L.stage;
            if (gtid == 0) {
                if (need_dx) {
                    mbar_expect_tx(full + s, (uint32_t)(2 * WT_IMG));
                    bulk_g2s(st + L.wt, S.wtimg + (size_t)p.j * tc_wt_chunk_floats(K, N), (uint32_t)(2 * WT_IMG), full + s);
                } else {
                    mbar_arrive(full + s);
                }
            }
#pragma unroll
            for (int i = 0; i < NXR; ++i) {
                const int q = qb + QS * i;
                const int c0 = CW * p.j + 4 * q;
                float4 a = x[i];
                if (!STAGE1) {
                    // a from the saved code; the code itself goes to the epilogue through the ring stage (it decodes
                    // the derivative w.r.t. the pre-activation and x-hat)
                    float4 m4 = make_float4(0.f, 0.f, 0.f, 0.f), xh = m4;
                    const float nanv = __int_as_float(0x7fc00000);
                    float4 code = make_float4(nanv, nanv, nanv, nanv);      // padding decodes to zeros
                    if (c0 < K && tile_row0(p.it) + r < B) { act_decode(x[i], sm_bna, Kp, c0, S.in.a.bn_mode, inv_keep, a, m4, xh); code = x[i]; }
                    else a = m4;
                    sts4(st + L.msk + sw128_chunk(r, q >> 3, q & 7, TCB_M), code);
                }
                float4 hi, lo;
                split_tf32x4(a, hi, lo);
                const float h4[4] = {hi.x, hi.y, hi.z, hi.w}, l4[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const uint32_t o2 = sw128_off(4 * q + e, r, ATR);
                    *reinterpret_cast<float*>(st + L.at + o2) = h4[e];
                    if (exact) *reinterpret_cast<float*>(st + L.at + AT_IMG + o2) = l4[e];
                }
            }
            fence_proxy_async();
            mbar_arrive(full + s);
            if (gtid == 0 && cnt < 40) trace(2 + cnt, tcode);
            // G / G^T of the tile are single-buffered and can only be rewritten when the previous tile's MMAs are done;
            // the tile's first chunk (which only needs a free ring stage) is therefore staged BEFORE G is rebuilt, so
            // the tensor core finds work the moment G arrives
            if (p.j == 0) build_g(p.it);
        };
        if constexpr (STAGE1) {
            chunk_pipeline<1, 2, NX, NXR>(my_tiles, nch, 0, issue_idx, issue_data, consume);
        } else {
            // saved codes through thread-private cp.async slots, two chunks in flight (exact wait_group waits; register
            // loads of a software pipeline all share one hardware scoreboard in these kernels, see tower_fwd_tc)
            constexpr int PD = 2;
            uint8_t* slots = sm + L.slots + gtid * 16;                  // + (slot * NXR + i) * TC_PT * 16
            ChunkPos pd, pc;
            pd.start(0, nch); pc = pd;
            int sd = 0, scs = 0;
            for (int jj = -PD;; ++jj) {
                if (jj >= 0 && pc.it >= my_tiles) break;
                if (jj >= 0) {
                    cp_async_wait<PD - 1>();
                    float4 x[NX];
#pragma unroll
                    for (int i = 0; i < NXR; ++i) x[i] = *reinterpret_cast<const float4*>(slots + (scs * NXR + i) * TC_PT * 16);
                    consume(pc, x);
                    pc.template next<1>(nch); scs ^= 1;
                }
                if (pd.it < my_tiles) {
                    const long long row = tile_row0(pd.it) + r;
#pragma unroll
                    for (int i = 0; i < NXR; ++i) {
                        const int c0 = CW * pd.j + 4 * (qb + QS * i);
                        const bool ok = row < B && c0 < K;
                        cp_async16(slots + (sd * NXR + i) * TC_PT * 16, S.in.a.a_post + (ok ? (size_t)row * K + c0 : 0), ok ? 16 : 0);
                    }
                }
                cp_async_commit();
                pd.template next<1>(nch); sd ^= 1;
            }
            cp_async_wait<0>();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (tid == 0) trace(127, tcode);
    if (warp == TC_MMA_WARP) tmem_dealloc(tmem_base, tmem_cols);
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static int st_N(const cfm_tower_t& t, int s) { return (int)(s == 1 ? t.h1 : s == 2 ? t.h2 : t.d_out); }
static int st_K(const cfm_tower_t& t, int s) { return (int)(s == 1 ? t.n_num + t.n_tables * t.emb_dim : s == 2 ? t.h1 : t.h2); }
static void tower_dims(const cfm_tower_t& t, int (&K)[3], int (&N)[3]) {
    for (int s = 1; s <= 3; ++s) { K[s - 1] = st_K(t, s); N[s - 1] = st_N(t, s); }
}

bool tc_fwd_supported(const cfm_tower_t& t, int s) {
    if (!t.wimg || !t.a1 || !t.a2) return false;
    const int K = st_K(t, s), N = st_N(t, s);
    if (N > 64 || K > 256) return false;
    if (s == 1) {
        if (t.n_tables > 0 && (t.emb_dim & 3)) return false;
    } else if (K & 3) return false;
    return tcf_smem(K, N, s == 1).nst >= 2;
}

bool tc_bwd_supported(const cfm_tower_t& t, int s, bool a_bn, bool need_dx) {
    if (!tc_fwd_supported(t, s)) return false;            // the forward of this layer saved what the backward reads
    const int K = st_K(t, s), N = st_N(t, s);
    if (N & 3) return false;
    if (s > 1 && K > 64) return false;
    (void)need_dx;
    return tcb_smem(K, N, s == 1, a_bn, s == 1 && t.xstash).nst >= 2;
}

int tc_prep_launch(const cfm_tower_t* towers, int n_towers, cudaStream_t stream) {
    PrepArgs a{};
    bool any = false;
    for (int i = 0; i < n_towers; ++i) {
        const cfm_tower_t& t = towers[i];
        PrepTower& P = a.t[i];
        P.W[0] = t.w1; P.W[1] = t.w2; P.W[2] = t.w3;
        tower_dims(t, P.K, P.N);
        P.KE = (int)(t.n_tables * t.emb_dim); P.n_num = (int)t.n_num;
        P.img = t.wimg;
        any |= t.wimg != nullptr;
    }
    if (!any) return CFM_OK;
    tc_prep_weights<<<dim3(64, (unsigned)n_towers), 256, 0, stream>>>(a);      // ~60 k elements per tower: latency-bound, spread wide
    CFM_LAUNCH_CHECK();
    return CFM_OK;
}

// Shares the SMs between the towers in proportion to their work per tile (chunks of the stage + a fixed part for the
// G / statistics / epilogue work every tile carries), at most one CTA per tile; returns the per-tower CTA counts.
static void tc_split_ctas(const int (&nch)[2], int n_towers, long long ntiles, int (&ctas)[2]) {
    const int sms = sm_count();
    ctas[0] = ctas[1] = 0;
    if (n_towers == 1) { ctas[0] = (int)std::min<long long>(ntiles, sms); return; }
    const double w0 = nch[0] + 1.5, w1 = nch[1] + 1.5;
    int c0 = (int)(sms * w0 / (w0 + w1) + 0.5);
    c0 = std::max(1, std::min(sms - 1, c0));
    ctas[0] = (int)std::min<long long>(ntiles, c0);
    ctas[1] = (int)std::min<long long>(ntiles, sms - c0);
}

int tc_fwd_launch(FwdArgs& a, const cfm_tower_t* towers, int n_towers, int s, int* ctas_out, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        CFM_CHECK_CUDA(cudaFuncSetAttribute(tower_fwd_tc<32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_MAX));
        CFM_CHECK_CUDA(cudaFuncSetAttribute(tower_fwd_tc<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_MAX));
        CFM_CHECK_CUDA(cudaFuncSetAttribute(tower_fwd_tc<32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_MAX));
        CFM_CHECK_CUDA(cudaFuncSetAttribute(tower_fwd_tc<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_MAX));
        attr_set = true;
    }
    size_t smem = 0;
    int nc = 32;
    int nchs[2] = {1, 1};
    for (int i = 0; i < n_towers; ++i) {
        const cfm_tower_t& t = towers[i];
        int K[3], N[3];
        tower_dims(t, K, N);
        const WImgLayout L = wimg_layout(K, N);
        nchs[i] = tc_nch(K[s - 1]);
        a.st[i].wimg = t.wimg + L.w[s - 1];
        a.st[i].a_out = s == 2 ? t.a1 : s == 3 ? t.a2 : nullptr;
        a.st[i].x_out = s == 1 ? t.xstash : nullptr;
        smem = std::max(smem, (size_t)tcf_smem(K[s - 1], N[s - 1], s == 1).total);
        if (tc_npad(N[s - 1]) > 32) nc = 64;
    }
    const long long ntiles = (a.B + TCF_M - 1) / TCF_M;
    int ctas[2];
    tc_split_ctas(nchs, n_towers, ntiles, ctas);
    ctas_out[0] = ctas[0]; ctas_out[1] = ctas[1];
    a.cta_split = ctas[0];
    const dim3 grid(ctas[0] + ctas[1]);
    if (s == 1) {
        if (nc == 64) tower_fwd_tc<64, true><<<grid, TC_THREADS, smem, stream>>>(a);
        else tower_fwd_tc<32, true><<<grid, TC_THREADS, smem, stream>>>(a);
    } else {
        if (nc == 64) tower_fwd_tc<64, false><<<grid, TC_THREADS, smem, stream>>>(a);
        else tower_fwd_tc<32, false><<<grid, TC_THREADS, smem, stream>>>(a);
    }
    CFM_LAUNCH_CHECK();
    return CFM_OK;
}

int tc_bwd_launch(BwdArgs& a, const cfm_tower_t* towers, int n_towers, int s, int* ctas_out, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        CFM_CHECK_CUDA(cudaFuncSetAttribute(tower_bwd_tc<true, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_MAX));
        CFM_CHECK_CUDA(cudaFuncSetAttribute(tower_bwd_tc<true, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_MAX));
        CFM_CHECK_CUDA(cudaFuncSetAttribute(tower_bwd_tc<false, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_MAX));
        CFM_CHECK_CUDA(cudaFuncSetAttribute(tower_bwd_tc<false, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_MAX));
        attr_set = true;
    }
    size_t smem = 0;
    int nchs[2] = {1, 1};
    for (int i = 0; i < n_towers; ++i) {
        const cfm_tower_t& t = towers[i];
        int K[3], N[3];
        tower_dims(t, K, N);
        const WImgLayout L = wimg_layout(K, N);
        nchs[i] = tcb_nch(K[s - 1]);
        a.st[i].wtimg = t.wimg + L.wt[s - 1];
        a.st[i].in.a.a_post = s == 2 ? t.a1 : s == 3 ? t.a2 : nullptr;
        a.st[i].x_in = s == 1 ? t.xstash : nullptr;
        smem = std::max(smem, (size_t)tcb_smem(K[s - 1], N[s - 1], s == 1, a.st[i].a_bn != 0, s == 1 && t.xstash).total);
    }
    const long long ntiles = (a.B + TCB_M - 1) / TCB_M;
    int ctas[2];
    tc_split_ctas(nchs, n_towers, ntiles, ctas);
    ctas_out[0] = ctas[0]; ctas_out[1] = ctas[1];
    a.cta_split = ctas[0];
    const dim3 grid(ctas[0] + ctas[1]);
    const int cw = tcb_cw(st_K(towers[0], s));
    if (s == 1) {
        if (cw == 64) tower_bwd_tc<true, 64><<<grid, TC_THREADS, smem, stream>>>(a);
        else tower_bwd_tc<true, 32><<<grid, TC_THREADS, smem, stream>>>(a);
    } else {
        if (cw == 64) tower_bwd_tc<false, 64><<<grid, TC_THREADS, smem, stream>>>(a);
        else tower_bwd_tc<false, 32><<<grid, TC_THREADS, smem, stream>>>(a);
    }
    CFM_LAUNCH_CHECK();
    return CFM_OK;
}

}  // namespace cfm

extern "C" int cfm_debug_set_trace(uint64_t* buf, int64_t code) {
    const int c = (int)code;
    CFM_CHECK_CUDA(cudaMemcpyToSymbol(cfm::g_trace, &buf, sizeof(buf)));
    CFM_CHECK_CUDA(cudaMemcpyToSymbol(cfm::g_trace_code, &c, sizeof(c)));
    return CFM_OK;
}

extern "C" int64_t cfm_tower_xstash_floats(const cfm_tower_t* t, int64_t B) {
    const int K = cfm::st_K(*t, 1);
    return ((B + 63) / 64) * cfm::tcb_nch(K) * 4096;
}

extern "C" int64_t cfm_tower_wimg_floats(const cfm_tower_t* t) {
    int K[3], N[3];
    cfm::tower_dims(*t, K, N);
    return cfm::wimg_layout(K, N).total;
}
