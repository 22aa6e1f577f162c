// Microbenchmark: random-row gather bandwidth on B200 (rows of RB bytes from 1M-row tables), three access paths.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_bw gather_bw.cu
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// (a) LDG.128: 8 lanes per 128-byte piece of a row; UNR independent loads per thread in flight
template <int UNR>
__global__ void gather_ldg(const float4* __restrict__ tab, const int* __restrict__ idx, long long n_items, int quads_per_row,
                           float* out) {
    // item = (row request, quad); consecutive threads -> consecutive quads of a row
    float acc = 0.f;
    const long long total = n_items * quads_per_row;
    for (long long base = (long long)blockIdx.x * blockDim.x * UNR + threadIdx.x; base < total; base += (long long)gridDim.x * blockDim.x * UNR) {
        float4 v[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const long long i = base + (long long)u * blockDim.x;
            v[u] = make_float4(0, 0, 0, 0);
            if (i < total) {
                const long long it = i / quads_per_row; const int q = (int)(i - it * quads_per_row);
                v[u] = __ldg(tab + (size_t)idx[it] * quads_per_row + q);
            }
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
    }
    if (acc == 123.456f) out[0] = acc;
}

// (b) cp.async.bulk: one bulk copy per row into shared memory, NB rows per batch per CTA, double-buffered
template <int RB>
__global__ void gather_bulk(const char* __restrict__ tab, const int* __restrict__ idx, long long n_items, float* out) {
    constexpr int NB = 256;                         // rows per batch (one per thread)
    extern __shared__ __align__(128) char sm[];
    __shared__ uint64_t bar[2];
    if (threadIdx.x == 0) {
        for (int b = 0; b < 2; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[b])), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    float acc = 0.f;
    const long long nb = (n_items + NB - 1) / NB;
    int it = 0;
    auto issue = [&](long long batch, int b) {
        const long long i = batch * NB + threadIdx.x;
        if (threadIdx.x == 0) {
            const long long cntb = min((long long)NB, n_items - batch * NB);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[b])), "r"((uint32_t)(cntb * RB)) : "memory");
        }
        __syncwarp();
        if (i < n_items) {
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_u32(sm + (size_t)b * NB * RB + (size_t)threadIdx.x * RB)),
                         "l"(tab + (size_t)idx[i] * RB), "r"(RB), "r"(smem_u32(&bar[b]))
                         : "memory");
        }
    };
    long long batch = blockIdx.x;
    if (batch < nb) issue(batch, 0);
    for (; batch < nb; batch += gridDim.x, ++it) {
        const int b = it & 1;
        const long long nxt = batch + gridDim.x;
        if (nxt < nb) issue(nxt, b ^ 1);
        // wait for this batch
        asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(&bar[b])), "r"((uint32_t)((it >> 1) & 1)) : "memory");
        acc += reinterpret_cast<float*>(sm + (size_t)b * NB * RB)[threadIdx.x];
        __syncthreads();
    }
    if (acc == 123.456f) out[0] = acc;
}

int main() {
    const long long ROWS = 1000000;
    const long long N = 4 * 65536 * 4;     // row requests
    for (int RB : {192, 32}) {
        const int qpr = RB / 16;
        char* tab; int* idx; float* out;
        CK(cudaMalloc(&tab, ROWS * RB * 4));       // 4 tables worth
        CK(cudaMemset(tab, 0, ROWS * RB * 4));
        CK(cudaMalloc(&idx, N * sizeof(int)));
        CK(cudaMalloc(&out, 4));
        std::vector<int> h(N);
        uint64_t s = 88172645463325252ull;
        for (long long i = 0; i < N; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (int)(s % (ROWS * 4)); }
        CK(cudaMemcpy(idx, h.data(), N * sizeof(int), cudaMemcpyHostToDevice));
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        auto report = [&](const char* name, float ms) { printf("RB=%3d %-28s %8.3f ms  %7.1f GB/s useful\n", RB, name, ms, (double)N * RB / ms / 1e6); };
        for (int blocks : {148 * 2, 148 * 4, 148 * 8}) {
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                gather_ldg<4><<<blocks, 256>>>((const float4*)tab, idx, N, qpr, out);
                cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (rep) { char nm[64]; snprintf(nm, 64, "ldg unr4 %d x256", blocks); report(nm, ms); }
            }
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                gather_ldg<8><<<blocks, 256>>>((const float4*)tab, idx, N, qpr, out);
                cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (rep) { char nm[64]; snprintf(nm, 64, "ldg unr8 %d x256", blocks); report(nm, ms); }
            }
        }
        for (int blocks : {148, 148 * 2}) {
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                if (RB == 192) { CK(cudaFuncSetAttribute(gather_bulk<192>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 256 * 192));
                    gather_bulk<192><<<blocks, 256, 2 * 256 * 192>>>(tab, idx, N, out); }
                else gather_bulk<32><<<blocks, 256, 2 * 256 * 32>>>(tab, idx, N, out);
                cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (rep) { char nm[64]; snprintf(nm, 64, "bulk %d x256", blocks); report(nm, ms); }
            }
        }
        CK(cudaGetLastError());
        cudaFree(tab); cudaFree(idx); cudaFree(out);
    }
    return 0;
}
