// Shared declarations of the tcgen05 tower stage kernels (tower_tc.cu) and their caller (tower.cu).
#pragma once
#include "tower_types.cuh"

namespace cfm {

// ---- weight images ------------------------------------------------------------------------------------------
// Per tower and Linear layer, rebuilt once per forward call by tc_prep_weights (the weights change every step):
//   forward image   : W [Npad rows][K columns] cut into chunks of 32 input columns; chunk j = hi image then lo
//                     image, each [Npad][32 floats] 128-byte swizzled (the K-major B operand of  H = X . W^T)
//   transposed image: W^T [K rows][N columns] cut into chunks of 64 rows (input columns); chunk j = hi then lo,
//                     each [64][nblk * 32 floats] (the K-major B operand of  dX[:, chunk j] = G . W[:, chunk j])
// Stage-1 input columns are in tile order [embedding columns | numeric columns] (gcol_stage1 maps back).
__host__ __device__ inline int tc_npad(int N) { return (N + 15) & ~15; }
__host__ __device__ inline int tc_nch(int K) { return (K + 31) >> 5; }
__host__ __device__ inline int tc_nblk(int N) { return (tc_npad(N) + 31) >> 5; }
__host__ __device__ inline int tc_w_chunk_floats(int N) { return 2 * tc_npad(N) * 32; }
// the backward walks the input columns in chunks of 64 (32 for layers with at most 32 inputs)
__host__ __device__ inline int tcb_cw(int K) { return K <= 32 ? 32 : 64; }
__host__ __device__ inline int tcb_nch(int K) { return (K + tcb_cw(K) - 1) / tcb_cw(K); }
__host__ __device__ inline int tc_wt_chunk_floats(int K, int N) { return 2 * tcb_cw(K) * tc_nblk(N) * 32; }
struct WImgLayout {
    int w[3], wt[3];      // float offsets of the forward / transposed images of layers 1..3
    int total;
};
__host__ __device__ inline WImgLayout wimg_layout(const int (&K)[3], const int (&N)[3]) {
    WImgLayout L;
    int o = 0;
    for (int s = 0; s < 3; ++s) {
        L.w[s] = o;  o += tc_nch(K[s]) * tc_w_chunk_floats(N[s]);
        L.wt[s] = o; o += tcb_nch(K[s]) * tc_wt_chunk_floats(K[s], N[s]);
    }
    L.total = o;
    return L;
}

// host entry points (tower_tc.cu)
bool tc_fwd_supported(const cfm_tower_t& t, int s);
bool tc_bwd_supported(const cfm_tower_t& t, int s, bool a_bn, bool need_dx);
int tc_prep_launch(const cfm_tower_t* towers, int n_towers, cudaStream_t stream);
int tc_fwd_launch(FwdArgs& a, const cfm_tower_t* towers, int n_towers, int s, int* ctas_out, cudaStream_t stream);
int tc_bwd_launch(BwdArgs& a, const cfm_tower_t* towers, int n_towers, int s, int* ctas_out, cudaStream_t stream);

}  // namespace cfm
