// Ceiling of the batched passes' gather traffic: every nonzero of the shared matrix pulls one 256-byte row
// (32 instances x 8 bytes) of the group's dense operand through the L2.  This microbenchmark does only that, with the
// launch shape of batched_rows_kernel (warp per matrix row, lane = instance, grid.y = instance groups, 6 CTAs/SM, indices
// prefetched 32 at a time and broadcast through shared memory), no epilogue streams, no matrix values:
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/rowgather_bench tools/rowgather_bench.cu
//   build/rowgather_bench      -> JSON lines: rows, nnz/row, operand rows V, groups, ms, GB/s through the L2
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ unsigned long long keep_policy() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ double ld_keep(const double *p, unsigned long long pol) {
    double v;
    asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
    return v;
}

template <int UNROLL>
__global__ void __launch_bounds__(256, 6) rowgather_kernel(int rows, int len, int V, const int *__restrict__ col, const double *__restrict__ slab,
                                                           double *__restrict__ out) {
    __shared__ int idx[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned long long keep = keep_policy();
    const double *g = slab + (size_t)blockIdx.y * V * 32 + lane;
    const int row0 = blockIdx.x * 32, row1 = min(rows, row0 + 32);
    double total = 0.0;
    for (int r = row0 + warp; r < row1; r += 8) {
        const size_t p0 = (size_t)r * len;
        double acc = 0.0;
        for (int k0 = 0; k0 < len; k0 += 32) {
            const int cnt = min(32, len - k0);
            __syncwarp();
            idx[warp][lane] = lane < cnt ? __ldg(col + p0 + k0 + lane) : 0;
            __syncwarp();
#pragma unroll UNROLL
            for (int t = 0; t < cnt; ++t) acc += ld_keep(g + (size_t)idx[warp][t] * 32, keep);
        }
        total += acc;
    }
    if (total == 1.2345e-300) out[blockIdx.x] = total;   // keep the loads alive
}

int main() {
    struct Case { int rows, len, V; const char *what; };
    const Case cases[] = {{200000, 10, 50000, "x-phase of configs[3]: rows of A^T gather Y"},
                          {50000, 40, 200000, "y-phase of configs[3]: rows of A gather X_hat"}};
    const int G = 8;
    for (const Case &c : cases) {
        const size_t nnz = (size_t)c.rows * c.len;
        std::vector<int> h(nnz);
        unsigned long long s = 0x9E3779B97F4A7C15ULL;
        for (size_t k = 0; k < nnz; ++k) { s = s * 6364136223846793005ULL + 1442695040888963407ULL; h[k] = (int)((s >> 33) % (unsigned long long)c.V); }
        int *col; double *slab, *out;
        CK(cudaMalloc(&col, nnz * 4)); CK(cudaMalloc(&slab, (size_t)G * c.V * 32 * 8)); CK(cudaMalloc(&out, 1 << 20));
        CK(cudaMemcpy(col, h.data(), nnz * 4, cudaMemcpyHostToDevice));
        CK(cudaMemset(slab, 0, (size_t)G * c.V * 32 * 8));
        const dim3 grid((c.rows + 31) / 32, G);
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        for (int i = 0; i < 3; ++i) rowgather_kernel<4><<<grid, 256>>>(c.rows, c.len, c.V, col, slab, out);
        CK(cudaEventRecord(e0));
        const int reps = 20;
        for (int i = 0; i < reps; ++i) rowgather_kernel<4><<<grid, 256>>>(c.rows, c.len, c.V, col, slab, out);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= reps;
        const double bytes = (double)nnz * G * 256.0;
        printf("{\"case\": \"%s\", \"rows\": %d, \"nnz_per_row\": %d, \"operand_rows\": %d, \"groups\": %d, \"slab_MB_per_group\": %.1f, \"ms\": %.4f, \"gather_GBps\": %.1f}\n",
               c.what, c.rows, c.len, c.V, G, c.V * 256.0 / 1e6, ms, bytes / (ms * 1e-3) / 1e9);
        CK(cudaFree(col)); CK(cudaFree(slab)); CK(cudaFree(out));
    }
    return 0;
}
