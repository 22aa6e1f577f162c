// Embedding-gradient scatter as a deterministic sorted-segment reduce (no atomics).
// replaces aten::embedding_dense_backward for the nn.Embedding tables of model.py:24-33.
//
// 1. keys = (table << idx_bits) | index for every (row, table) pair, values = row          (1 kernel)
// 2. stable LSD radix sort of (key, row) over the minimal number of key bits (cub)          (library sort)
// 3. one thread group per segment head walks its run of equal keys in row order, sums the
//    [E]-wide slices of dx_emb and writes the dense gradient row once                         (1 kernel)
// Equal keys keep increasing row order (stable sort), so the floating-point sum order is fixed.
#include "common.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <algorithm>

namespace cfm {

struct TablePtrs { float* p[CFM_MAX_TABLES]; long long rows[CFM_MAX_TABLES]; };

__global__ void emb_make_keys(const long long* __restrict__ x_cat, long long B, int n_tab, int idx_bits,
                              unsigned long long* __restrict__ keys, int* __restrict__ vals, TablePtrs tp) {
    const long long n = B * n_tab;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        // i enumerates x_cat in memory order (coalesced read); slot t*B + r keeps rows increasing per table
        long long r = i / n_tab;
        int t = (int)(i - r * n_tab);
        long long idx = x_cat[i];
        if (idx < 0 || idx >= tp.rows[t]) idx = 0;   // forward already flagged the error
        keys[(long long)t * B + r] = ((unsigned long long)t << idx_bits) | (unsigned long long)idx;
        vals[(long long)t * B + r] = (int)r;
    }
}

// thread (p, q): if sorted position p starts a run, sum float4 slice q of every row in the run
__global__ void emb_segment_reduce(const unsigned long long* __restrict__ keys, const int* __restrict__ vals,
                                   long long n, int n_tab, int E, int idx_bits, const float* __restrict__ dx,
                                   TablePtrs tp) {
    const int E4 = E >> 2;
    const long long total = n * E4;
    const int KE = n_tab * E;
    const unsigned long long mask = (1ull << idx_bits) - 1;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long p = i / E4;
        int q = (int)(i - p * E4);
        unsigned long long key = keys[p];
        if (p > 0 && keys[p - 1] == key) continue;
        int t = (int)(key >> idx_bits);
        long long idx = (long long)(key & mask);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (long long s = p; s < n && keys[s] == key; ++s) {
            const float4 g = __ldg(reinterpret_cast<const float4*>(dx + (size_t)vals[s] * KE + t * E) + q);
            acc.x += g.x; acc.y += g.y; acc.z += g.z; acc.w += g.w;
        }
        reinterpret_cast<float4*>(tp.p[t] + (size_t)idx * E)[q] = acc;
    }
}

// scalar variant for E % 4 != 0
__global__ void emb_segment_reduce_scalar(const unsigned long long* __restrict__ keys, const int* __restrict__ vals,
                                          long long n, int n_tab, int E, int idx_bits, const float* __restrict__ dx,
                                          TablePtrs tp) {
    const long long total = n * E;
    const int KE = n_tab * E;
    const unsigned long long mask = (1ull << idx_bits) - 1;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long p = i / E;
        int e = (int)(i - p * E);
        unsigned long long key = keys[p];
        if (p > 0 && keys[p - 1] == key) continue;
        int t = (int)(key >> idx_bits);
        long long idx = (long long)(key & mask);
        float acc = 0.f;
        for (long long s = p; s < n && keys[s] == key; ++s) acc += dx[(size_t)vals[s] * KE + t * E + e];
        tp.p[t][(size_t)idx * E + e] = acc;
    }
}

__global__ void emb_rezero(const unsigned long long* __restrict__ keys, long long n, int E, int idx_bits, TablePtrs tp) {
    const long long total = n * E;
    const unsigned long long mask = (1ull << idx_bits) - 1;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long p = i / E;
        int e = (int)(i - p * E);
        unsigned long long key = keys[p];
        if (p > 0 && keys[p - 1] == key) continue;
        tp.p[(int)(key >> idx_bits)][(size_t)(key & mask) * E + e] = 0.f;
    }
}

static int bits_for(long long n) { int b = 1; while ((1ll << b) < n) ++b; return b; }

static int fill_tables(TablePtrs& tp, float* const* grad_tables, const int64_t* table_rows, int64_t n_tables,
                       int* idx_bits, int* key_bits) {
    CFM_REQUIRE(n_tables >= 1 && n_tables <= CFM_MAX_TABLES, CFM_ERR_INVALID, "n_tables outside [1,%d]", CFM_MAX_TABLES);
    long long max_rows = 1;
    for (int i = 0; i < CFM_MAX_TABLES; ++i) {
        tp.p[i] = i < n_tables ? grad_tables[i] : nullptr;
        tp.rows[i] = i < n_tables ? table_rows[i] : 0;
        if (i < n_tables) {
            CFM_REQUIRE(grad_tables[i] && table_rows[i] >= 1, CFM_ERR_INVALID, "bad gradient table %d", i);
            max_rows = std::max<long long>(max_rows, table_rows[i]);
        }
    }
    *idx_bits = bits_for(max_rows);
    *key_bits = *idx_bits + bits_for(n_tables);
    CFM_REQUIRE(*key_bits <= 62, CFM_ERR_UNSUPPORTED, "tables too large for the 64-bit sort key");
    return CFM_OK;
}

}  // namespace cfm

using namespace cfm;

extern "C" int64_t cfm_emb_grad_tmp_bytes(int64_t n_tables, int64_t B) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                    (const int*)nullptr, (int*)nullptr, (int)(n_tables * B), 0, 64, (cudaStream_t)0);
    return (int64_t)bytes + 256;
}

extern "C" int cfm_emb_grad_segment_reduce(const int64_t* x_cat, const float* dx_emb, int64_t B, int64_t n_tables,
                                           int64_t emb_dim, float* const* grad_tables, const int64_t* table_rows,
                                           int64_t* keys_tmp, int32_t* vals_tmp, int64_t* keys_sorted,
                                           int32_t* vals_sorted, void* sort_tmp, int64_t sort_tmp_bytes,
                                           void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CFM_REQUIRE(x_cat && dx_emb && grad_tables && table_rows && keys_tmp && vals_tmp && keys_sorted && vals_sorted &&
                    sort_tmp, CFM_ERR_INVALID, "null pointer");
    CFM_REQUIRE(B >= 1 && emb_dim >= 1 && n_tables * B < (1ll << 31), CFM_ERR_INVALID, "bad sizes");
    TablePtrs tp;
    int idx_bits, key_bits;
    int rc = fill_tables(tp, grad_tables, table_rows, n_tables, &idx_bits, &key_bits);
    if (rc) return rc;
    const long long n = n_tables * B;
    const int grid = (int)std::min<long long>((n + 255) / 256, 148 * 8);
    ProfScope prof(PROF_EMB, stream);
    emb_make_keys<<<grid, 256, 0, stream>>>((const long long*)x_cat, B, (int)n_tables, idx_bits,
                                            (unsigned long long*)keys_tmp, vals_tmp, tp);
    CFM_LAUNCH_CHECK();
    size_t bytes = (size_t)sort_tmp_bytes;
    CFM_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(sort_tmp, bytes, (const unsigned long long*)keys_tmp,
                                                   (unsigned long long*)keys_sorted, (const int*)vals_tmp,
                                                   vals_sorted, (int)n, 0, key_bits, stream));
    if ((emb_dim & 3) == 0) {
        const long long total = n * (emb_dim / 4);
        emb_segment_reduce<<<(int)std::min<long long>((total + 255) / 256, 148 * 16), 256, 0, stream>>>(
            (const unsigned long long*)keys_sorted, vals_sorted, n, (int)n_tables, (int)emb_dim, idx_bits, dx_emb, tp);
    } else {
        const long long total = n * emb_dim;
        emb_segment_reduce_scalar<<<(int)std::min<long long>((total + 255) / 256, 148 * 16), 256, 0, stream>>>(
            (const unsigned long long*)keys_sorted, vals_sorted, n, (int)n_tables, (int)emb_dim, idx_bits, dx_emb, tp);
    }
    CFM_LAUNCH_CHECK();
    return CFM_OK;
}

extern "C" int cfm_emb_grad_rezero(float* const* grad_tables, const int64_t* table_rows, int64_t n_tables,
                                   int64_t emb_dim, const int64_t* keys_sorted, int64_t n_items, void* stream_) {
    CFM_REQUIRE(grad_tables && table_rows && keys_sorted, CFM_ERR_INVALID, "null pointer");
    if (n_items == 0) return CFM_OK;
    TablePtrs tp;
    int idx_bits, key_bits;
    int rc = fill_tables(tp, grad_tables, table_rows, n_tables, &idx_bits, &key_bits);
    if (rc) return rc;
    const long long total = n_items * emb_dim;
    emb_rezero<<<(int)std::min<long long>((total + 255) / 256, 148 * 16), 256, 0, (cudaStream_t)stream_>>>(
        (const unsigned long long*)keys_sorted, n_items, (int)emb_dim, idx_bits, tp);
    CFM_LAUNCH_CHECK();
    return CFM_OK;
}
