// Descriptors of the tower stage kernels (tower.cu).
#pragma once
#include "common.cuh"

namespace cfm {

constexpr int TM = 64;    // rows per tile (maximum; the widest stage runs 32-row tiles so two CTAs fit one SM)
constexpr int NT = 256;   // threads per CTA
constexpr float BN_EPS = 1e-5f;
constexpr int MAX_SMEM = 227 * 1024;

// ------------------------------------------------------------------------------------------
// descriptors (plain structs passed by value as kernel parameters)
// ------------------------------------------------------------------------------------------
struct GatherSrc {
    int n_num, n_tab, E;
    const float* x_num;
    const long long* x_cat;
    const float* tab[CFM_MAX_TABLES];
    long long tab_rows[CFM_MAX_TABLES];
};

struct ActSrc {           // a = dropout(relu(bn(h)))
    const float* h;       // [B,K] raw Linear output of the previous stage
    int bn_mode;          // 0 none, 1 batch stats (mean, istd), 2 running stats (mean, var)
    const float *mean, *var_or_istd, *gamma, *beta;
    DropCtx drop;
    const float* a_post;  // tcgen05 backward: a itself as saved by the forward stage (nullable)
};

struct InputDesc {        // how a stage's [TM, K] input tile is built
    int stage;            // 1: gather, >1: activation of previous stage
    int K;
    GatherSrc g;
    ActSrc a;
};

// Warp-level tensor-core tiling (mma.sync m16n8k8, TF32 operands split 3x for fp32-class accuracy) of a
// [TM=64, n] output: 8 warps = 2 row halves (32 rows, two m16 tiles) x 4 column groups of 8*NI columns, in passes
// of up to 128 columns.
struct MmaPlan {
    int passes;
    int wrows;            // rows of the [n, K]-style operand kept in smem (n rounded up to 32)
};
__host__ __device__ inline MmaPlan make_plan(int n) {
    MmaPlan p;
    p.passes = ceil_div(n, 128);
    p.wrows = (n + 31) & ~31;
    return p;
}
__host__ __device__ inline int pass_cols(int n, int pass) { return min(128, n - pass * 128); }
// n-tiles of 8 columns per warp so that `wc` column groups of warps cover `cols` (<= 128) columns
__host__ __device__ inline int pass_ni(int cols, int wc) { return cols <= 8 * wc ? 1 : cols <= 16 * wc ? 2 : 4; }
__host__ __device__ inline int ceil8(int k) { return (k + 7) & ~7; }
// leading dimension for "transposed" fragment loads (lanes walk rows with t, columns with g): ld == 8 (mod 32)
__host__ __device__ inline int pad_ld_t(int k) { int ld = (ceil8(k) + 31) & ~31; return ld + 8; }

struct FwdStage {
    InputDesc in;
    int N;
    const float* W;       // [N,K]
    const float* bias;    // [N]
    float* hout;          // [B,N]
    float* stat_part;     // nullable: per-CTA (mean[N], M2[N], count) partials
    const float* wimg;    // tcgen05 path: forward weight image of this layer (tower_tc.cuh)
    float* a_out;         // tcgen05 path, stage > 1: where the producers save the input activation a (nullable)
    float* x_out;         // tcgen05 path, stage 1: stash of the gathered input tiles (nullable; tower_tc.cu)
};
struct FwdArgs {
    FwdStage st[2];
    long long B;
    int* err;
    int tm;               // rows per tile: 64 (2 row groups x 4 column groups of warps) or 32 (1 x 8)
    int exact;            // tcgen05 path: 1 = three (hi, lo) passes, 0 = single TF32 pass
    int cta_split;        // tcgen05 path: 1-D grid, CTAs [0, cta_split) work on tower 0, the rest on tower 1
};

struct BwdStage {
    InputDesc in;         // rebuilds the stage's input tile A
    int N;
    const float* W;       // [N,K]
    int a_bn;             // BN precedes the input activation: keep x-hat tile and emit (sum dy, sum dy*xhat)
    const float* gin;     // [B,N] incoming gradient (g_out for stage 3, dy_s otherwise)
    int g_mode;           // 0 as is; 1 train BN: gamma*istd*(dy - c1 - xhat*c2); 2 eval BN: dy*gamma*rsqrt(var+eps)
    const float* hs;      // [B,N] raw output of this stage (g_mode 1)
    const float *g_mean, *g_var_or_istd, *g_gamma, *g_c1, *g_c2;
    float* dW_part;       // per-CTA [N, K+1] (last column = bias gradient), global column order
    float* dy_out;        // stage>1: [B,K]
    float* sum_part;      // a_bn: per-CTA [2K]
    float* dx_emb;        // stage 1, nullable: [B, n_tab*E]
    float* dx_num;        // stage 1, nullable: [B, n_num]
    int need_dx;          // run the dX GEMM
    const float* wtimg;   // tcgen05 path: transposed weight image of this layer (tower_tc.cuh)
    const float* x_in;    // tcgen05 path, stage 1: the forward's stash of input tiles (nullable: gather again)
};
struct BwdArgs {
    BwdStage st[2];
    long long B;
    int tm;               // rows per tile, see FwdArgs
    int exact;            // see FwdArgs
    int cta_split;        // see FwdArgs
};


// smem column c' of a stage-1 tile -> column of the torch concat layout [numeric | emb_0 | emb_1 ...]
// (the tile keeps embeddings first so every gathered row lands 16-byte aligned)
__device__ __forceinline__ int gcol_stage1(int c, int KE, int n_num) { return c < KE ? n_num + c : c - KE; }

}  // namespace cfm
