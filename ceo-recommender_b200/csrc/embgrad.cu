// Embedding-gradient scatter as a deterministic sorted-segment reduce (no atomics).
// replaces aten::embedding_dense_backward for the nn.Embedding tables of model.py:24-33.
//
// 1. keys = (table << idx_bits) | index for every (row, table) pair, values = row          (1 kernel)
// 2. stable LSD radix sort of (key, row) over the minimal number of key bits (cub)          (library sort)
// 3. every run of equal keys is cut into chunks of SEG_CHUNK positions counted from the run's first position; one
//    thread per (chunk, 16-byte slice) sums its chunk in position order; a run that fits one chunk is written
//    straight to the dense gradient row, longer runs leave per-chunk partials that a second kernel adds in a fixed
//    order (one warp per run: lane l takes chunks l, l + 32, ..., then a shuffle tree)            (2 kernels)
// Equal keys keep increasing row order (stable sort) and the chunk grid hangs on the run itself, so the
// floating-point sum order is fixed whatever the launch geometry - and a low-cardinality column (runs of tens of
// thousands of rows: data.py:120-126 has tables of 2-4 classes) is reduced by hundreds of threads instead of one.
#include "common.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <algorithm>

namespace cfm {

struct TablePtrs { float* p[CFM_MAX_TABLES]; long long rows[CFM_MAX_TABLES]; };

// ------------------------------------------------------------------------------------------
// two-level fixed-order segment reduce shared by the single-tower, joint and peer variants
// ------------------------------------------------------------------------------------------
constexpr int SEG_CHUNK = 64;
constexpr int SEG_MAX_WIDTH = 256;            // floats per destination row slice the partial scratch is sized for

// starts[p] = first sorted position of the run p belongs to: run heads marked with their own position, then an inclusive
// max-scan (cub).  Computed once per sort, beside it (it only needs the sorted keys).
__global__ void seg_mark_heads(const unsigned long long* __restrict__ keys, long long n, int* __restrict__ starts) {
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x)
        starts[p] = (p > 0 && keys[p - 1] == keys[p]) ? 0 : (int)p;
}
// Is position p (local to a range that begins at global position `base`) the first position of a chunk of its run?
__device__ __forceinline__ bool seg_chunk_head(const unsigned long long* __restrict__ keys, const int* __restrict__ starts,
                                               long long base, long long p, unsigned long long key, long long& start) {
    if (p == 0 || keys[p - 1] != key) { start = p; return true; }
    start = (long long)starts[p] - base;
    return ((p - start) % SEG_CHUNK) == 0;
}
// Partial slot of the chunk at position p of a run starting at `start`.  A window of SEG_CHUNK positions holds at most
// two heads of multi-chunk runs: one of a run that began before the window, one of a run that begins inside it.
__device__ __forceinline__ long long seg_slot(long long p, long long start) {
    const long long w = p / SEG_CHUNK;
    return 2 * w + (start >= w * SEG_CHUNK ? 1 : 0);
}

// level 1: thread (p, q) sums slice q of the chunk that starts at p
// general item: any run length (chunk heads only), bisection for the chunk end, four rows in flight
template <int VEC, class Acc>
__device__ __forceinline__ void seg_item(const unsigned long long* __restrict__ keys, const int* __restrict__ vals, long long n,
                                         const Acc& acc, float* __restrict__ part, int* __restrict__ any_long,
                                         const int* __restrict__ starts, long long base, int W, long long p, int q,
                                         unsigned long long key, unsigned long long kprev, unsigned long long knext) {
    long long start;
    if (kprev != key) start = p;
    else {
        start = (long long)starts[p] - base;
        if (((p - start) % SEG_CHUNK) != 0) return;              // not the first position of a chunk of its run
    }
    // end of the chunk: a full chunk unless the run stops inside it (then a 6-step bisection finds where), so the
    // summation loop below carries no key test and its loads are independent of one another
    long long end = min(n, p + SEG_CHUNK);
    if (knext != key) end = p + 1;
    else if (keys[end - 1] != key) {
        long long lo = p, hi = end - 1;                    // keys[lo] == key, keys[hi] != key
        while (hi - lo > 1) {
            const long long mid = (lo + hi) >> 1;
            if (keys[mid] == key) lo = mid; else hi = mid;
        }
        end = hi;
    }
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    long long s = p;
    for (; s + 4 <= end; s += 4) {                         // four rows in flight, added in position order
        const float* s0 = acc.src(key, vals[s], q);
        const float* s1 = acc.src(key, vals[s + 1], q);
        const float* s2 = acc.src(key, vals[s + 2], q);
        const float* s3 = acc.src(key, vals[s + 3], q);
        if (VEC == 4) {
            const float4 g0 = *reinterpret_cast<const float4*>(s0), g1 = *reinterpret_cast<const float4*>(s1);
            const float4 g2 = *reinterpret_cast<const float4*>(s2), g3 = *reinterpret_cast<const float4*>(s3);
            a.x += g0.x; a.y += g0.y; a.z += g0.z; a.w += g0.w;
            a.x += g1.x; a.y += g1.y; a.z += g1.z; a.w += g1.w;
            a.x += g2.x; a.y += g2.y; a.z += g2.z; a.w += g2.w;
            a.x += g3.x; a.y += g3.y; a.z += g3.z; a.w += g3.w;
        } else {
            const float g0 = *s0, g1 = *s1, g2 = *s2, g3 = *s3;
            a.x += g0; a.x += g1; a.x += g2; a.x += g3;
        }
    }
    for (; s < end; ++s) {
        const float* src = acc.src(key, vals[s], q);
        if (VEC == 4) {
            const float4 g = *reinterpret_cast<const float4*>(src);
            a.x += g.x; a.y += g.y; a.z += g.z; a.w += g.w;
        } else {
            a.x += *src;
        }
    }
    // the run is this one chunk: it starts here and ends inside the chunk (or exactly at its end)
    const bool single = start == p && (end < p + SEG_CHUNK || p + SEG_CHUNK >= n || keys[p + SEG_CHUNK] != key);
    float* dst = single ? acc.dst(key, q) : part + (seg_slot(p, start) * W + q) * VEC;
    if (VEC == 4) *reinterpret_cast<float4*>(dst) = a;
    else *dst = a.x;
    if (!single && start == p && q == 0) *any_long = 1;
}
template <int VEC, class Acc>
__global__ void seg_reduce_chunks(const unsigned long long* __restrict__ keys, const int* __restrict__ vals, long long n, Acc acc,
                                  float* __restrict__ part, int* __restrict__ any_long, const int* __restrict__ starts,
                                  long long base) {
    const int W = acc.slices();
    const long long total = n * W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / W;
        const int q = (int)(i - p * W);
        // the three neighbouring keys are requested together: with large tables most runs have ONE position (previous
        // and next key both differ) and the item needs no further key traffic - one load latency instead of a chain of
        // eight (head test, end test, six bisection steps)
        const unsigned long long key = keys[p];
        const unsigned long long kprev = p > 0 ? keys[p - 1] : ~key, knext = p + 1 < n ? keys[p + 1] : ~key;
        seg_item<VEC, Acc>(keys, vals, n, acc, part, any_long, starts, base, W, p, q, key, kprev, knext);
    }
}
// level 2: one WARP per sorted position; if it is the first position of a multi-chunk run, lane l adds the run's chunk
// partials l, l + 32, ... in increasing order and the 32 lane sums are added in lane order (a fixed tree), slice by slice
template <int VEC, class Acc>
__global__ void seg_reduce_long_runs(const unsigned long long* __restrict__ keys, long long n, Acc acc,
                                     const float* __restrict__ part, const int* __restrict__ any_long) {
    if (*any_long == 0) return;
    const int W = acc.slices(), lane = threadIdx.x & 31;
    const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long p = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; p < n; p += warps) {
        const unsigned long long key = keys[p];
        if (p > 0 && keys[p - 1] == key) continue;
        if (p + SEG_CHUNK >= n || keys[p + SEG_CHUNK] != key) continue;
        long long lo = p + SEG_CHUNK, hi = n;                  // keys[lo] == key; first position past the run in (lo, hi]
        while (hi - lo > 1) {
            const long long mid = (lo + hi) >> 1;
            if (keys[mid] == key) lo = mid; else hi = mid;
        }
        const long long nchunks = (hi - p + SEG_CHUNK - 1) / SEG_CHUNK;
        for (int q = 0; q < W; ++q) {
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            for (long long c = lane; c < nchunks; c += 32) {
                const float* src = part + (seg_slot(p + c * SEG_CHUNK, p) * W + q) * VEC;
                if (VEC == 4) {
                    const float4 g = *reinterpret_cast<const float4*>(src);
                    a.x += g.x; a.y += g.y; a.z += g.z; a.w += g.w;
                } else {
                    a.x += *src;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                a.x += __shfl_down_sync(FULL, a.x, o);
                if (VEC == 4) {
                    a.y += __shfl_down_sync(FULL, a.y, o);
                    a.z += __shfl_down_sync(FULL, a.z, o);
                    a.w += __shfl_down_sync(FULL, a.w, o);
                }
            }
            if (lane == 0) {
                float* dst = acc.dst(key, q);
                if (VEC == 4) *reinterpret_cast<float4*>(dst) = a;
                else *dst = a.x;
            }
        }
    }
}

// sources / destinations of the dense case: rows of dx [*, n_tab * E], tables t0 .. of `tp`
template <int VEC>
struct DenseAcc {
    const float* dx;
    TablePtrs tp;
    int KE, E, idx_bits, t0;
    __device__ int slices() const { return E / VEC; }
    __device__ const float* src(unsigned long long key, int val, int q) const {
        return dx + (size_t)val * KE + ((int)(key >> idx_bits) - t0) * E + q * VEC;
    }
    __device__ float* dst(unsigned long long key, int q) const {
        return tp.p[(int)(key >> idx_bits) - t0] + (size_t)(key & ((1ull << idx_bits) - 1)) * E + q * VEC;
    }
};

// scratch behind cub's temporary storage: chunk partials + the "some run is long" flag
static size_t cub_sort_bytes(long long n) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                    (const int*)nullptr, (int*)nullptr, (int)n, 0, 64, (cudaStream_t)0);
    size_t scan = 0;                                   // the run-start scan reuses the region after the sort
    cub::DeviceScan::InclusiveScan(nullptr, scan, (int*)nullptr, (int*)nullptr, cub::Max(), (int)n, (cudaStream_t)0);
    bytes = std::max(bytes, scan);
    return (bytes + 255) & ~(size_t)255;
}
static size_t seg_part_bytes(long long n) { return (size_t)(n / SEG_CHUNK + 2) * 2 * SEG_MAX_WIDTH * sizeof(float); }
struct SegScratch { float* part; int* flag; int* starts; size_t cub_bytes; };
static size_t seg_scratch_bytes(long long n) { return cub_sort_bytes(n) + seg_part_bytes(n) + 256 + (size_t)n * sizeof(int) + 256; }
static int seg_scratch(void* sort_tmp, int64_t sort_tmp_bytes, long long n, SegScratch* out) {
    const size_t cb = cub_sort_bytes(n), pb = seg_part_bytes(n);
    CFM_REQUIRE((size_t)sort_tmp_bytes >= seg_scratch_bytes(n), CFM_ERR_INVALID,
                "sort scratch of %lld bytes is too small (cfm_emb_grad_tmp_bytes gives the size)", (long long)sort_tmp_bytes);
    out->cub_bytes = cb;
    out->part = reinterpret_cast<float*>(static_cast<char*>(sort_tmp) + cb);
    out->flag = reinterpret_cast<int*>(static_cast<char*>(sort_tmp) + cb + pb);
    out->starts = reinterpret_cast<int*>(static_cast<char*>(sort_tmp) + cb + pb + 256);
    return CFM_OK;
}
// run starts of the freshly sorted keys (uses cub's temporary storage again: the sort is done with it)
static int seg_run_starts(const unsigned long long* keys_sorted, long long n, void* sort_tmp, const SegScratch& sc, cudaStream_t stream) {
    seg_mark_heads<<<(int)std::min<long long>((n + 255) / 256, 148 * 8), 256, 0, stream>>>(keys_sorted, n, sc.starts);
    CFM_LAUNCH_CHECK();
    size_t need = 0;
    cub::DeviceScan::InclusiveScan(nullptr, need, sc.starts, sc.starts, cub::Max(), (int)n, stream);
    CFM_REQUIRE(need <= sc.cub_bytes, CFM_ERR_INVALID, "scan scratch larger than the sort scratch");
    size_t bytes = sc.cub_bytes;
    CFM_CHECK_CUDA(cub::DeviceScan::InclusiveScan(sort_tmp, bytes, sc.starts, sc.starts, cub::Max(), (int)n, stream));
    return CFM_OK;
}
template <int VEC, class Acc>
static int seg_reduce_launch(const unsigned long long* keys, const int* vals, long long n, const Acc& acc, int width,
                             const SegScratch& sc, long long base, cudaStream_t stream) {
    CFM_REQUIRE(width <= SEG_MAX_WIDTH, CFM_ERR_UNSUPPORTED, "embedding slices wider than %d floats are not supported", SEG_MAX_WIDTH);
    CFM_CHECK_CUDA(cudaMemsetAsync(sc.flag, 0, sizeof(int), stream));
    const long long total = n * (width / VEC);
    const int grid = (int)std::min<long long>((total + 255) / 256, 148 * 16);
    seg_reduce_chunks<VEC, Acc><<<grid, 256, 0, stream>>>(keys, vals, n, acc, sc.part, sc.flag, sc.starts + base, base);
    CFM_LAUNCH_CHECK();
    const int grid2 = (int)std::min<long long>((n * 32 + 255) / 256, 148 * 8);     // normally a no-op (no long runs): keep its launch cheap
    seg_reduce_long_runs<VEC, Acc><<<grid2, 256, 0, stream>>>(keys, n, acc, sc.part, sc.flag);
    CFM_LAUNCH_CHECK();
    return CFM_OK;
}

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
    return (int64_t)seg_scratch_bytes(n_tables * B);
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
    SegScratch sc;
    rc = seg_scratch(sort_tmp, sort_tmp_bytes, n, &sc);
    if (rc) return rc;
    size_t bytes = sc.cub_bytes;
    CFM_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(sort_tmp, bytes, (const unsigned long long*)keys_tmp,
                                                   (unsigned long long*)keys_sorted, (const int*)vals_tmp,
                                                   vals_sorted, (int)n, 0, key_bits, stream));
    const unsigned long long* ks = (const unsigned long long*)keys_sorted;
    rc = seg_run_starts(ks, n, sort_tmp, sc, stream);
    if (rc) return rc;
    if ((emb_dim & 3) == 0) {
        DenseAcc<4> acc{dx_emb, tp, (int)(n_tables * emb_dim), (int)emb_dim, idx_bits, 0};
        return seg_reduce_launch<4>(ks, vals_sorted, n, acc, (int)emb_dim, sc, 0, stream);
    }
    DenseAcc<1> acc{dx_emb, tp, (int)(n_tables * emb_dim), (int)emb_dim, idx_bits, 0};
    return seg_reduce_launch<1>(ks, vals_sorted, n, acc, (int)emb_dim, sc, 0, stream);
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

// =============================================================================================
// Table-sharded embeddings over NVLink peer memory (single node, one process per GPU).
// Every table -- or every column slice ("piece") of a wide table -- lives on ONE rank; the others map it (CUDA IPC)
// and read rows straight through NVLink:
//   * emb_gather_rows      pulls the rows a batch needs from local/peer tables into a local stash [K][B][E]
//   * emb_make_keys_peer / emb_segment_reduce_peer: the owner reduces the per-pair gradient rows of EVERY rank's
//     batch, reading the peers' dx_emb buffers in place (no staging copy, no all-to-all).
// Order inside a run of equal keys is (rank, row) ascending == the order of the concatenated global batch,
// so the result is bitwise the replicated data-parallel one.
// =============================================================================================
namespace cfm {

struct GatherPlan {
    const float* table[CFM_MAX_SLOTS];      // [k * pieces + piece]: base of table k as mapped from that piece's owner
    long long rows[CFM_MAX_TABLES];
};

// unit i -> (pair j = b*K + k in x_cat memory order, vector q of the row); U independent loads in flight per thread
// IDX: unsigned (32-bit index arithmetic: B*K*E/VEC < 2^31, the usual case; 64-bit division costs ~10x more and this
// kernel issues four per 16 bytes moved) or long long
template <int VEC, typename IDX>
__global__ void emb_gather_rows(const long long* __restrict__ x_cat, long long B_, int K_, int E, int pieces, int w,
                                GatherPlan gp, float* __restrict__ stash, int* __restrict__ err) {
    const IDX B = (IDX)B_, K = (IDX)K_;
    const IDX EV = (IDX)(E / VEC);
    const IDX total = B * K * EV;
    constexpr int U = 4;
    const IDX stride = (IDX)gridDim.x * blockDim.x;
    for (IDX i = (IDX)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride * U) {
        float4 v4[U];
        float v1[U];
        long long dst[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const IDX iu = i + (IDX)u * stride;
            dst[u] = -1;
            if (iu < total && iu >= i) {                         // (iu >= i: no wrap-around in the 32-bit variant)
                const IDX j = iu / EV;
                const int c = (int)(iu - j * EV) * VEC;          // first column of this vector
                const IDX b = j / K;
                const int k = (int)(j - b * K);
                long long idx = x_cat[j];
                if (idx < 0 || idx >= gp.rows[k]) {
                    if (c == 0) atomicOr(err, 1);
                    idx = 0;
                }
                const float* src = gp.table[k * pieces + c / w] + (size_t)idx * E + c;
                if (VEC == 4) v4[u] = *reinterpret_cast<const float4*>(src);
                else v1[u] = *src;
                dst[u] = ((long long)k * (long long)B + (long long)b) * E + c;
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (dst[u] >= 0) {
                if (VEC == 4) *reinterpret_cast<float4*>(stash + dst[u]) = v4[u];
                else stash[dst[u]] = v1[u];
            }
        }
    }
}

struct PeerPlan {
    const long long* x_cat[CFM_MAX_TABLES][CFM_MAX_PEERS];
    const float* dx[CFM_MAX_TABLES][CFM_MAX_PEERS];
    int n_cols[CFM_MAX_TABLES];
    int col[CFM_MAX_TABLES];
    int col0[CFM_MAX_TABLES];
    int E[CFM_MAX_TABLES];
};

// slot (j*W + r)*B + b: owned-slice-major, then rank, then row -> a stable sort keeps (rank, row) order per key
__global__ void emb_make_keys_peer(PeerPlan pp, int n_owned, int W, long long B, int idx_bits, TablePtrs tp,
                                   unsigned long long* __restrict__ keys, int* __restrict__ vals) {
    const long long n = (long long)n_owned * W * B;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long jr = i / B;
        const long long b = i - jr * B;
        const int j = (int)(jr / W), r = (int)(jr - (long long)j * W);
        long long idx = pp.x_cat[j][r][b * pp.n_cols[j] + pp.col[j]];
        if (idx < 0 || idx >= tp.rows[j]) idx = 0;       // the gather already flagged the error
        keys[i] = ((unsigned long long)j << idx_bits) | (unsigned long long)idx;
        vals[i] = (int)((long long)r * B + b);
    }
}

// sources / destinations of the owner-side reduce: item value = rank * B + row, gradient rows read in place from the
// rank's (peer-mapped) buffer, destination = columns [col0, col0 + w) of the owned table
template <int VEC>
struct PeerAcc {
    PeerPlan pp;
    TablePtrs tp;
    int B, E, w, idx_bits;
    __device__ int slices() const { return w / VEC; }
    __device__ const float* src(unsigned long long key, int val, int q) const {
        const int j = (int)(key >> idx_bits);
        const int r = val / B;                                  // n_owned * n_peers * B < 2^31, so B fits an int
        const long long b = val - r * B;
        return pp.dx[j][r] + (size_t)b * (pp.n_cols[j] * E) + pp.col[j] * E + pp.col0[j] + q * VEC;
    }
    __device__ float* dst(unsigned long long key, int q) const {
        const int j = (int)(key >> idx_bits);
        return tp.p[j] + (size_t)(key & ((1ull << idx_bits) - 1)) * E + pp.col0[j] + q * VEC;
    }
};

// zero columns [col0, col0 + w) of every row a sorted key names
template <int VEC>
__global__ void emb_rezero_peer(const unsigned long long* __restrict__ keys, long long n, int E, int w, int idx_bits,
                                PeerPlan pp, TablePtrs tp) {
    const int WV = w / VEC;
    const long long total = n * WV;
    const unsigned long long mask = (1ull << idx_bits) - 1;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / WV;
        const int q = (int)(i - p * WV);
        const unsigned long long key = keys[p];
        if (p > 0 && keys[p - 1] == key) continue;
        const int j = (int)(key >> idx_bits);
        float* dst = tp.p[j] + (size_t)(key & mask) * E + pp.col0[j] + q * VEC;
        if (VEC == 4) *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
        else *dst = 0.f;
    }
}

// one plan over the owned slices of ALL groups (slice ids are global, group-major)
static int fill_peer_plan(PeerPlan& pp, TablePtrs& tp, const cfm_peer_group_t* groups, int64_t n_groups, int64_t n_peers,
                          bool need_peers, int* idx_bits, int* key_bits, int* total_owned) {
    CFM_REQUIRE(groups && n_groups >= 1 && n_groups <= CFM_MAX_GROUPS, CFM_ERR_INVALID, "n_groups outside [1,%d]",
                CFM_MAX_GROUPS);
    CFM_REQUIRE(n_peers >= 1 && n_peers <= CFM_MAX_PEERS, CFM_ERR_INVALID, "n_peers outside [1,%d]", CFM_MAX_PEERS);
    float* grads[CFM_MAX_TABLES];
    int64_t rows[CFM_MAX_TABLES];
    for (int j = 0; j < CFM_MAX_TABLES; ++j) {
        pp.n_cols[j] = pp.col[j] = pp.col0[j] = pp.E[j] = 0;
        for (int r = 0; r < CFM_MAX_PEERS; ++r) { pp.x_cat[j][r] = nullptr; pp.dx[j][r] = nullptr; }
    }
    int j = 0;
    for (int g = 0; g < n_groups; ++g) {
        const cfm_peer_group_t& G = groups[g];
        CFM_REQUIRE(G.owned && G.n_owned >= 1 && j + G.n_owned <= CFM_MAX_TABLES, CFM_ERR_INVALID,
                    "group %d: bad slice count / more than %d owned slices", g, CFM_MAX_TABLES);
        CFM_REQUIRE(G.emb_dim >= 1 && G.width >= 1 && G.width <= G.emb_dim, CFM_ERR_INVALID, "group %d: bad emb_dim / width", g);
        for (int o = 0; o < G.n_owned; ++o, ++j) {
            const cfm_peer_table_t& t = G.owned[o];
            CFM_REQUIRE(t.n_cols >= 1 && t.col >= 0 && t.col < t.n_cols && t.grad && t.rows >= 1 && t.col0 >= 0 &&
                            t.col0 + G.width <= G.emb_dim, CFM_ERR_INVALID, "group %d: bad owned slice %d", g, o);
            pp.n_cols[j] = (int)t.n_cols; pp.col[j] = (int)t.col; pp.col0[j] = (int)t.col0; pp.E[j] = (int)G.emb_dim;
            grads[j] = t.grad; rows[j] = t.rows;
            for (int r = 0; need_peers && r < n_peers; ++r) {
                CFM_REQUIRE(t.x_cat[r] && t.dx_emb[r], CFM_ERR_INVALID, "group %d slice %d: null peer buffer %d", g, o, r);
                pp.x_cat[j][r] = (const long long*)t.x_cat[r];
                pp.dx[j][r] = t.dx_emb[r];
            }
        }
    }
    *total_owned = j;
    return fill_tables(tp, grads, rows, j, idx_bits, key_bits);
}

}  // namespace cfm

extern "C" int cfm_enable_peer_access(int64_t peer_device) {
    int dev = 0, can = 0;
    CFM_CHECK_CUDA(cudaGetDevice(&dev));
    if (peer_device == dev) return CFM_OK;
    CFM_CHECK_CUDA(cudaDeviceCanAccessPeer(&can, dev, (int)peer_device));
    CFM_REQUIRE(can, CFM_ERR_UNSUPPORTED, "device %d cannot access peer device %d", dev, (int)peer_device);
    cudaError_t e = cudaDeviceEnablePeerAccess((int)peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { (void)cudaGetLastError(); return CFM_OK; }
    CFM_CHECK_CUDA(e);
    return CFM_OK;
}

extern "C" int cfm_emb_gather_rows(const int64_t* x_cat, int64_t B, int64_t n_tables, int64_t emb_dim, int64_t pieces,
                                   const float* const* tables, const int64_t* table_rows, float* stash,
                                   int32_t* err_flag, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CFM_REQUIRE(x_cat && tables && table_rows && stash && err_flag, CFM_ERR_INVALID, "null pointer");
    CFM_REQUIRE(B >= 1 && emb_dim >= 1 && n_tables >= 1 && n_tables <= CFM_MAX_TABLES, CFM_ERR_INVALID, "bad sizes");
    CFM_REQUIRE(pieces >= 1 && emb_dim % pieces == 0 && n_tables * pieces <= CFM_MAX_SLOTS, CFM_ERR_INVALID,
                "pieces must divide emb_dim and n_tables*pieces <= %d", CFM_MAX_SLOTS);
    GatherPlan gp;
    for (int i = 0; i < CFM_MAX_SLOTS; ++i) {
        gp.table[i] = i < n_tables * pieces ? tables[i] : nullptr;
        if (i < n_tables * pieces) CFM_REQUIRE(tables[i], CFM_ERR_INVALID, "null table slice %d", i);
    }
    for (int i = 0; i < CFM_MAX_TABLES; ++i) {
        gp.rows[i] = i < n_tables ? table_rows[i] : 0;
        if (i < n_tables) CFM_REQUIRE(table_rows[i] >= 1, CFM_ERR_INVALID, "bad table %d", i);
    }
    const int w = (int)(emb_dim / pieces);
    const bool vec = (w & 3) == 0;
    const long long total = B * n_tables * (vec ? emb_dim / 4 : emb_dim);
    const int grid = (int)std::max<long long>(1, std::min<long long>((total + 1023) / 1024, (long long)sm_count() * 8));
    ProfScope prof(PROF_EMB, stream);
    // 32-bit index arithmetic unless the unit count (plus the unrolled look-ahead) could overflow it
    const bool small = total + 4ll * grid * 256 < (1ll << 31);
    const long long* xc = (const long long*)x_cat;
    if (vec && small) emb_gather_rows<4, unsigned><<<grid, 256, 0, stream>>>(xc, B, (int)n_tables, (int)emb_dim, (int)pieces, w, gp, stash, err_flag);
    else if (vec) emb_gather_rows<4, long long><<<grid, 256, 0, stream>>>(xc, B, (int)n_tables, (int)emb_dim, (int)pieces, w, gp, stash, err_flag);
    else if (small) emb_gather_rows<1, unsigned><<<grid, 256, 0, stream>>>(xc, B, (int)n_tables, (int)emb_dim, (int)pieces, w, gp, stash, err_flag);
    else emb_gather_rows<1, long long><<<grid, 256, 0, stream>>>(xc, B, (int)n_tables, (int)emb_dim, (int)pieces, w, gp, stash, err_flag);
    CFM_LAUNCH_CHECK();
    return CFM_OK;
}

namespace cfm {
// Second stream of the embedding-gradient reduce (one per device): the two towers' segment reductions write disjoint
// tables and read disjoint key ranges, so the second group runs beside the first; under stream capture the fork / join
// events become graph edges.
struct EmbSide { cudaStream_t stream = nullptr; cudaEvent_t fork = nullptr, join = nullptr; };
static EmbSide* emb_side() {
    static EmbSide side[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    EmbSide& S = side[dev];
    if (!S.stream) {
        if (cudaStreamCreateWithFlags(&S.stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        cudaEventCreateWithFlags(&S.fork, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&S.join, cudaEventDisableTiming);
    }
    return &S;
}
}  // namespace cfm

extern "C" int cfm_emb_grad_peer_reduce(const cfm_peer_group_t* groups, int64_t n_groups, int64_t n_peers, int64_t B,
                                        int64_t phase, int64_t* keys_tmp, int32_t* vals_tmp, int64_t* keys_sorted,
                                        int32_t* vals_sorted, void* sort_tmp, int64_t sort_tmp_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CFM_REQUIRE(keys_tmp && vals_tmp && keys_sorted && vals_sorted && sort_tmp, CFM_ERR_INVALID, "null pointer");
    CFM_REQUIRE(phase >= 0 && phase <= 2, CFM_ERR_INVALID, "phase must be 0 (all), 1 (keys + sort) or 2 (reduce)");
    PeerPlan pp;
    TablePtrs tp;
    int idx_bits, key_bits, n_owned;
    int rc = fill_peer_plan(pp, tp, groups, n_groups, n_peers, true, &idx_bits, &key_bits, &n_owned);
    if (rc) return rc;
    CFM_REQUIRE(B >= 1 && (long long)n_owned * n_peers * B < (1ll << 31), CFM_ERR_INVALID, "bad sizes");
    const long long n = (long long)n_owned * n_peers * B;
    ProfScope prof(PROF_EMB, stream);
    SegScratch sc;
    rc = seg_scratch(sort_tmp, sort_tmp_bytes, n, &sc);
    if (rc) return rc;
    if (phase != 2) {
        emb_make_keys_peer<<<(int)std::min<long long>((n + 255) / 256, 148 * 8), 256, 0, stream>>>(
            pp, n_owned, (int)n_peers, B, idx_bits, tp, (unsigned long long*)keys_tmp, vals_tmp);
        CFM_LAUNCH_CHECK();
        size_t bytes = sc.cub_bytes;
        CFM_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(sort_tmp, bytes, (const unsigned long long*)keys_tmp,
                                                       (unsigned long long*)keys_sorted, (const int*)vals_tmp,
                                                       vals_sorted, (int)n, 0, key_bits, stream));
        rc = seg_run_starts((const unsigned long long*)keys_sorted, n, sort_tmp, sc, stream);
        if (rc) return rc;
    }
    if (phase == 1) return CFM_OK;
    // the second group reduces beside the first (own slice of the partial scratch, own flag, second stream)
    size_t part_off[CFM_MAX_GROUPS + 1] = {0};
    for (int g = 0; g < n_groups; ++g)
        part_off[g + 1] = part_off[g] + (size_t)(groups[g].n_owned * n_peers * B / SEG_CHUNK + 2) * 2 * (size_t)groups[g].width;
    EmbSide* side = n_groups == 2 && part_off[n_groups] * sizeof(float) <= seg_part_bytes(n) ? emb_side() : nullptr;
    if (side) {
        CFM_CHECK_CUDA(cudaEventRecord(side->fork, stream));
        CFM_CHECK_CUDA(cudaStreamWaitEvent(side->stream, side->fork, 0));
    }
    long long off = 0;
    for (int g = 0; g < n_groups; ++g) {
        const cfm_peer_group_t& G = groups[g];
        const long long ng = G.n_owned * n_peers * B;
        const unsigned long long* ks = (const unsigned long long*)keys_sorted + off;
        SegScratch scg = sc;
        cudaStream_t st = stream;
        if (side) { scg.part = sc.part + part_off[g]; scg.flag = sc.flag + g; if (g == 1) st = side->stream; }
        if (ng > 0) {
            if ((G.width & 3) == 0 && (G.emb_dim & 3) == 0) {
                PeerAcc<4> acc{pp, tp, (int)B, (int)G.emb_dim, (int)G.width, idx_bits};
                rc = seg_reduce_launch<4>(ks, vals_sorted + off, ng, acc, (int)G.width, scg, off, st);
            } else {
                PeerAcc<1> acc{pp, tp, (int)B, (int)G.emb_dim, (int)G.width, idx_bits};
                rc = seg_reduce_launch<1>(ks, vals_sorted + off, ng, acc, (int)G.width, scg, off, st);
            }
            if (rc) return rc;
        }
        off += ng;
    }
    if (side) {
        CFM_CHECK_CUDA(cudaEventRecord(side->join, side->stream));
        CFM_CHECK_CUDA(cudaStreamWaitEvent(stream, side->join, 0));
    }
    return CFM_OK;
}

extern "C" int cfm_emb_grad_peer_rezero(const cfm_peer_group_t* groups, int64_t n_groups, int64_t n_peers, int64_t B,
                                        const int64_t* keys_sorted, void* stream_) {
    CFM_REQUIRE(keys_sorted, CFM_ERR_INVALID, "null pointer");
    PeerPlan pp;
    TablePtrs tp;
    int idx_bits, key_bits, n_owned;
    int rc = fill_peer_plan(pp, tp, groups, n_groups, n_peers, false, &idx_bits, &key_bits, &n_owned);
    if (rc) return rc;
    long long off = 0;
    for (int g = 0; g < n_groups; ++g) {
        const cfm_peer_group_t& G = groups[g];
        const long long ng = G.n_owned * n_peers * B;
        if ((G.width & 3) == 0 && (G.emb_dim & 3) == 0) {
            const long long total = ng * (G.width / 4);
            emb_rezero_peer<4><<<(int)std::min<long long>((total + 255) / 256, 148 * 16), 256, 0, (cudaStream_t)stream_>>>(
                (const unsigned long long*)keys_sorted + off, ng, (int)G.emb_dim, (int)G.width, idx_bits, pp, tp);
        } else {
            const long long total = ng * G.width;
            emb_rezero_peer<1><<<(int)std::min<long long>((total + 255) / 256, 148 * 16), 256, 0, (cudaStream_t)stream_>>>(
                (const unsigned long long*)keys_sorted + off, ng, (int)G.emb_dim, (int)G.width, idx_bits, pp, tp);
        }
        CFM_LAUNCH_CHECK();
        off += ng;
    }
    return CFM_OK;
}

// ---- CUDA IPC plumbing for the peer mappings (export on the owner, open on the accessor's own device) ----
typedef int (*MemGetAddressRangeFn)(unsigned long long*, size_t*, unsigned long long);

extern "C" int cfm_ipc_export(const void* ptr, uint8_t* handle_out, int64_t* offset_out) {
    CFM_REQUIRE(ptr && handle_out && offset_out, CFM_ERR_INVALID, "null pointer");
    static MemGetAddressRangeFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        CFM_CHECK_CUDA(cudaGetDriverEntryPoint("cuMemGetAddressRange", &p, cudaEnableDefault, &q));
        CFM_REQUIRE(q == cudaDriverEntryPointSuccess && p, CFM_ERR_CUDA, "cuMemGetAddressRange entry point not available");
        fn = (MemGetAddressRangeFn)p;
    }
    unsigned long long base = 0;
    size_t size = 0;
    int rc = fn(&base, &size, (unsigned long long)(uintptr_t)ptr);
    CFM_REQUIRE(rc == 0, CFM_ERR_CUDA, "cuMemGetAddressRange failed (%d)", rc);
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, (void*)(uintptr_t)base);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        CFM_REQUIRE(false, CFM_ERR_CUDA,
                    "cudaIpcGetMemHandle: %s (the buffer must come from cudaMalloc: expandable segments / "
                    "cudaMallocAsync pools cannot be exported)", cudaGetErrorString(e));
    }
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(handle_out, &h, sizeof(h));
    *offset_out = (int64_t)((unsigned long long)(uintptr_t)ptr - base);
    return CFM_OK;
}

extern "C" int cfm_ipc_open(const uint8_t* handle, void** base_out) {
    CFM_REQUIRE(handle && base_out, CFM_ERR_INVALID, "null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    // opened from the ACCESSING device's context: peer access to the exporting device is enabled lazily
    CFM_CHECK_CUDA(cudaIpcOpenMemHandle(base_out, h, cudaIpcMemLazyEnablePeerAccess));
    return CFM_OK;
}

extern "C" int cfm_ipc_close(void* base) {
    CFM_REQUIRE(base, CFM_ERR_INVALID, "null pointer");
    CFM_CHECK_CUDA(cudaIpcCloseMemHandle(base));
    return CFM_OK;
}

// =============================================================================================
// Joint reduce for the towers of one step: ONE key build and ONE radix sort over the (table, index) pairs of every
// tower (the sort passes are latency-bound at these sizes, so two half-size sorts cost almost twice one), then one
// segment-reduce launch per tower on its contiguous key range (tables are numbered tower-major, so after the sort
// tower g owns positions [B * sum_{h<g} K_h, B * sum_{h<=g} K_h)).
// =============================================================================================
namespace cfm {

struct JointKeys {
    const long long* x_cat[CFM_MAX_GROUPS];
    int n_tab[CFM_MAX_GROUPS];
    int t0[CFM_MAX_GROUPS];                 // first global table id of the group
    long long rows[CFM_MAX_TABLES];         // by global table id
    int n_groups;
};

__global__ void emb_make_keys_joint(JointKeys jk, long long B, int idx_bits, unsigned long long* __restrict__ keys,
                                    int* __restrict__ vals) {
    for (int g = 0; g < jk.n_groups; ++g) {
        const int K = jk.n_tab[g], t0 = jk.t0[g];
        const long long n = B * K;
        const long long* __restrict__ x = jk.x_cat[g];
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
            const long long r = i / K;
            const int t = (int)(i - r * K);
            long long idx = x[i];
            if (idx < 0 || idx >= jk.rows[t0 + t]) idx = 0;   // forward already flagged the error
            keys[(long long)(t0 + t) * B + r] = ((unsigned long long)(t0 + t) << idx_bits) | (unsigned long long)idx;
            vals[(long long)(t0 + t) * B + r] = (int)r;
        }
    }
}

template <int VEC>
__global__ void emb_rezero_range(const unsigned long long* __restrict__ keys, long long n, int E, int idx_bits, int t0,
                                 TablePtrs tp) {
    const int EV = E / VEC;
    const long long total = n * EV;
    const unsigned long long mask = (1ull << idx_bits) - 1;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / EV;
        const int q = (int)(i - p * EV);
        const unsigned long long key = keys[p];
        if (p > 0 && keys[p - 1] == key) continue;
        float* dst = tp.p[(int)(key >> idx_bits) - t0] + (size_t)(key & mask) * E + q * VEC;
        if (VEC == 4) *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
        else *dst = 0.f;
    }
}

// validates the groups; fills per-group table pointers, the joint key plan and the key widths
static int joint_plan(const cfm_emb_group_t* groups, int64_t n_groups, bool need_x, bool need_dx, JointKeys& jk,
                      TablePtrs (&tps)[CFM_MAX_GROUPS], int* idx_bits, int* key_bits, long long* total_tables) {
    CFM_REQUIRE(groups && n_groups >= 1 && n_groups <= CFM_MAX_GROUPS, CFM_ERR_INVALID, "n_groups outside [1,%d]",
                CFM_MAX_GROUPS);
    long long tt = 0, max_rows = 1;
    jk.n_groups = (int)n_groups;
    for (int g = 0; g < CFM_MAX_GROUPS; ++g) { jk.x_cat[g] = nullptr; jk.n_tab[g] = 0; jk.t0[g] = 0; }
    for (int i = 0; i < CFM_MAX_TABLES; ++i) jk.rows[i] = 0;
    for (int g = 0; g < n_groups; ++g) {
        const cfm_emb_group_t& G = groups[g];
        CFM_REQUIRE(G.n_tables >= 1 && G.emb_dim >= 1 && tt + G.n_tables <= CFM_MAX_TABLES, CFM_ERR_INVALID,
                    "group %d: bad table count / more than %d tables in total", g, CFM_MAX_TABLES);
        CFM_REQUIRE((!need_x || G.x_cat) && (!need_dx || G.dx_emb), CFM_ERR_INVALID, "group %d: null input", g);
        jk.x_cat[g] = (const long long*)G.x_cat; jk.n_tab[g] = (int)G.n_tables; jk.t0[g] = (int)tt;
        for (int i = 0; i < CFM_MAX_TABLES; ++i) {
            tps[g].p[i] = i < G.n_tables ? G.grad_tables[i] : nullptr;
            tps[g].rows[i] = i < G.n_tables ? G.table_rows[i] : 0;
            if (i < G.n_tables) {
                CFM_REQUIRE(G.grad_tables[i] && G.table_rows[i] >= 1, CFM_ERR_INVALID, "group %d: bad table %d", g, i);
                jk.rows[tt + i] = G.table_rows[i];
                max_rows = std::max<long long>(max_rows, G.table_rows[i]);
            }
        }
        tt += G.n_tables;
    }
    *idx_bits = bits_for(max_rows);
    *key_bits = *idx_bits + bits_for(tt);
    *total_tables = tt;
    CFM_REQUIRE(*key_bits <= 62, CFM_ERR_UNSUPPORTED, "tables too large for the 64-bit sort key");
    return CFM_OK;
}

}  // namespace cfm

extern "C" int cfm_emb_grad_joint_reduce(const cfm_emb_group_t* groups, int64_t n_groups, int64_t B, int64_t phase,
                                         int64_t* keys_tmp, int32_t* vals_tmp, int64_t* keys_sorted, int32_t* vals_sorted,
                                         void* sort_tmp, int64_t sort_tmp_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CFM_REQUIRE(keys_tmp && vals_tmp && keys_sorted && vals_sorted && sort_tmp, CFM_ERR_INVALID, "null pointer");
    JointKeys jk;
    TablePtrs tps[CFM_MAX_GROUPS];
    int idx_bits, key_bits;
    long long tt;
    CFM_REQUIRE(phase >= 0 && phase <= 2, CFM_ERR_INVALID, "phase must be 0, 1 or 2");
    int rc = joint_plan(groups, n_groups, phase != 2, phase != 1, jk, tps, &idx_bits, &key_bits, &tt);
    if (rc) return rc;
    CFM_REQUIRE(B >= 1 && tt * B < (1ll << 31), CFM_ERR_INVALID, "bad sizes");
    const long long n = tt * B;
    ProfScope prof(PROF_EMB, stream);
    SegScratch sc;
    rc = seg_scratch(sort_tmp, sort_tmp_bytes, n, &sc);
    if (rc) return rc;
    if (phase != 2) {
        emb_make_keys_joint<<<(int)std::min<long long>((n + 255) / 256, 148 * 8), 256, 0, stream>>>(
            jk, B, idx_bits, (unsigned long long*)keys_tmp, vals_tmp);
        CFM_LAUNCH_CHECK();
        size_t bytes = sc.cub_bytes;
        CFM_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(sort_tmp, bytes, (const unsigned long long*)keys_tmp,
                                                       (unsigned long long*)keys_sorted, (const int*)vals_tmp,
                                                       vals_sorted, (int)n, 0, key_bits, stream));
        rc = seg_run_starts((const unsigned long long*)keys_sorted, n, sort_tmp, sc, stream);
        if (rc) return rc;
    }
    if (phase == 1) return CFM_OK;
    // group g > 0 gets its own slice of the partial scratch and its own flag, and runs on the second stream
    size_t part_off[CFM_MAX_GROUPS + 1] = {0};
    for (int g = 0; g < n_groups; ++g)
        part_off[g + 1] = part_off[g] + (size_t)(groups[g].n_tables * B / SEG_CHUNK + 2) * 2 * (size_t)groups[g].emb_dim;
    EmbSide* side = n_groups == 2 && part_off[n_groups] * sizeof(float) <= seg_part_bytes(n) ? emb_side() : nullptr;
    if (side) {
        CFM_CHECK_CUDA(cudaEventRecord(side->fork, stream));
        CFM_CHECK_CUDA(cudaStreamWaitEvent(side->stream, side->fork, 0));
    }
    for (int g = 0; g < n_groups; ++g) {
        const cfm_emb_group_t& G = groups[g];
        const long long off = (long long)jk.t0[g] * B, ng = G.n_tables * B;
        const unsigned long long* ks = (const unsigned long long*)keys_sorted + off;
        SegScratch scg = sc;
        cudaStream_t st = stream;
        if (side) { scg.part = sc.part + part_off[g]; scg.flag = sc.flag + g; if (g == 1) st = side->stream; }
        if ((G.emb_dim & 3) == 0) {
            DenseAcc<4> acc{G.dx_emb, tps[g], (int)(G.n_tables * G.emb_dim), (int)G.emb_dim, idx_bits, jk.t0[g]};
            rc = seg_reduce_launch<4>(ks, vals_sorted + off, ng, acc, (int)G.emb_dim, scg, off, st);
        } else {
            DenseAcc<1> acc{G.dx_emb, tps[g], (int)(G.n_tables * G.emb_dim), (int)G.emb_dim, idx_bits, jk.t0[g]};
            rc = seg_reduce_launch<1>(ks, vals_sorted + off, ng, acc, (int)G.emb_dim, scg, off, st);
        }
        if (rc) return rc;
    }
    if (side) {
        CFM_CHECK_CUDA(cudaEventRecord(side->join, side->stream));
        CFM_CHECK_CUDA(cudaStreamWaitEvent(stream, side->join, 0));
    }
    return CFM_OK;
}

extern "C" int cfm_emb_grad_joint_rezero(const cfm_emb_group_t* groups, int64_t n_groups, int64_t B,
                                         const int64_t* keys_sorted, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CFM_REQUIRE(keys_sorted, CFM_ERR_INVALID, "null pointer");
    JointKeys jk;
    TablePtrs tps[CFM_MAX_GROUPS];
    int idx_bits, key_bits;
    long long tt;
    int rc = joint_plan(groups, n_groups, false, false, jk, tps, &idx_bits, &key_bits, &tt);
    if (rc) return rc;
    CFM_REQUIRE(B >= 1, CFM_ERR_INVALID, "bad sizes");
    ProfScope prof(PROF_EMB, stream);
    for (int g = 0; g < n_groups; ++g) {
        const cfm_emb_group_t& G = groups[g];
        const long long off = (long long)jk.t0[g] * B, ng = G.n_tables * B;
        const unsigned long long* ks = (const unsigned long long*)keys_sorted + off;
        if ((G.emb_dim & 3) == 0) {
            const long long total = ng * (G.emb_dim / 4);
            emb_rezero_range<4><<<(int)std::min<long long>((total + 255) / 256, 148 * 16), 256, 0, stream>>>(
                ks, ng, (int)G.emb_dim, idx_bits, jk.t0[g], tps[g]);
        } else {
            const long long total = ng * G.emb_dim;
            emb_rezero_range<1><<<(int)std::min<long long>((total + 255) / 256, 148 * 16), 256, 0, stream>>>(
                ks, ng, (int)G.emb_dim, idx_bits, jk.t0[g], tps[g]);
        }
        CFM_LAUNCH_CHECK();
    }
    return CFM_OK;
}
