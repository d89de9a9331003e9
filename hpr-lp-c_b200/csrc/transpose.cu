// transpose.cu -- device construction of A^T from A (setup path; SURVEY.md 8f rank 1: the step before the hot
// path).  Produces exactly the entry order of the reference's host counting sort (CSR_transpose_host,
// src/utils.cu:203-232): A^T rows in column order, entries of one A^T row ordered by original row index.
// That order is what a STABLE sort of the nonzeros by column index yields, so:
//   row id per nonzero (binary search in rowPtr) -> stable LSD radix sort of (col, nnz index) ->
//   gather row ids / values through the permutation -> row pointers by binary search in the sorted keys.
// cub::DeviceRadixSort is used for the sort (CUDA toolkit header library, setup only -- never on the
// iteration path).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <stdexcept>
#include <vector>

#include "engine.h"
#include "kernels.cuh"

namespace hpr {
namespace {

__global__ void nnz_row_ids_kernel(const int *rowPtr, int rows, int nnz, int *rowid, int *idx) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nnz) return;
    int lo = 0, hi = rows;   // first r with rowPtr[r+1] > k
    while (lo < hi) {
        const int mid = lo + ((hi - lo) >> 1);
        if (rowPtr[mid + 1] > k) hi = mid; else lo = mid + 1;
    }
    rowid[k] = lo;
    idx[k] = k;
}

__global__ void permute_kernel(const int *perm, const int *rowid, const double *val, int nnz, int *tcol, double *tval) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nnz) return;
    const int k = perm[j];
    tcol[j] = rowid[k];
    tval[j] = val[k];
}

__global__ void col_ptr_kernel(const int *sorted_cols, int nnz, int cols, int *trp) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > cols) return;
    int lo = 0, hi = nnz;    // first j with sorted_cols[j] >= c
    while (lo < hi) {
        const int mid = lo + ((hi - lo) >> 1);
        if (sorted_cols[mid] >= c) hi = mid; else lo = mid + 1;
    }
    trp[c] = lo;
}


// ---- column bands (engine.cu, Engine::build_bands) ------------------------------------------------------------------
// counts[b * (rows + 1) + r] = entries of row r whose column lies in band b (band = col / band_cols); one warp per row.
constexpr int kMaxBands = 64;
__global__ void band_count_kernel(const int *rowPtr, const int *col, int rows, int band_cols, int n_bands, int *counts) {
    __shared__ int cnt[8][kMaxBands];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int r = blockIdx.x * 8 + w;
    for (int b = lane; b < n_bands; b += 32) cnt[w][b] = 0;
    __syncwarp();
    if (r < rows) {
        for (int k = rowPtr[r] + lane; k < rowPtr[r + 1]; k += 32) atomicAdd(&cnt[w][col[k] / band_cols], 1);
    }
    __syncwarp();
    if (r < rows)
        for (int b = lane; b < n_bands; b += 32) counts[(size_t)b * (rows + 1) + r] = cnt[w][b];
}
// Stable split: within (row, band) the entries keep their order in the row, so a banded pass adds the same products as the
// plain pass, band by band.  One warp per row; 32 entries at a time, ranks by __match_any_sync.
__global__ void band_fill_kernel(const int *rowPtr, const int *col, const double *val, int rows, int band_cols, int n_bands,
                                 const int *band_rowPtr /* [n_bands][rows + 1] */, int *const *bcol, double *const *bval) {
    __shared__ int off[8][kMaxBands];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int r = blockIdx.x * 8 + w;
    for (int b = lane; b < n_bands; b += 32) off[w][b] = (r < rows) ? band_rowPtr[(size_t)b * (rows + 1) + r] : 0;
    __syncwarp();
    if (r >= rows) return;
    const int p0 = rowPtr[r], p1 = rowPtr[r + 1];
    for (int k0 = p0; k0 < p1; k0 += 32) {
        const int k = k0 + lane;
        const bool on = k < p1;
        const unsigned act = __ballot_sync(0xffffffffu, on);
        if (on) {
            const int c = col[k];
            const int b = c / band_cols;
            const unsigned same = __match_any_sync(act, b);
            const int rank = __popc(same & ((1u << lane) - 1u));
            const int pos = off[w][b] + rank;
            bcol[b][pos] = c;
            bval[b][pos] = val[k];
            __syncwarp(act);
            if (rank == 0) off[w][b] += __popc(same);
        }
        __syncwarp();
    }
}

}  // namespace

void device_transpose_csr(int rows, int cols, int nnz, const int *d_rowPtr, const int *d_col, const double *d_val,
                          int *d_trp, int *d_tcol, double *d_tval, cudaStream_t st) {
    int *rowid = nullptr, *idx = nullptr, *keys_out = nullptr, *perm = nullptr;
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    // stream-ordered temporaries from the engines' private memory pool (cached across solves, no synchronous cudaMalloc/cudaFree)
    int dev = 0;
    cudaGetDevice(&dev);
    auto pool_get = [&](size_t bytes) { return pool_alloc_raw(bytes, dev, st); };
    rowid = static_cast<int *>(pool_get(sizeof(int) * (size_t)nnz));
    idx = static_cast<int *>(pool_get(sizeof(int) * (size_t)nnz));
    keys_out = static_cast<int *>(pool_get(sizeof(int) * (size_t)nnz));
    perm = static_cast<int *>(pool_get(sizeof(int) * (size_t)nnz));
    const int T = 256;
    nnz_row_ids_kernel<<<(nnz + T - 1) / T, T, 0, st>>>(d_rowPtr, rows, nnz, rowid, idx);
    int bits = 1;
    while (bits < 31 && (1LL << bits) < (long long)cols) ++bits;
    HPR_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_col, keys_out, idx, perm, nnz, 0, bits, st));
    tmp = pool_get(std::max<size_t>(tmp_bytes, 16));
    HPR_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, d_col, keys_out, idx, perm, nnz, 0, bits, st));
    permute_kernel<<<(nnz + T - 1) / T, T, 0, st>>>(perm, rowid, d_val, nnz, d_tcol, d_tval);
    col_ptr_kernel<<<(cols + 1 + T - 1) / T, T, 0, st>>>(keys_out, nnz, cols, d_trp);
    cudaFreeAsync(rowid, st); cudaFreeAsync(idx, st); cudaFreeAsync(keys_out, st); cudaFreeAsync(perm, st); cudaFreeAsync(tmp, st);
}

// Splits a device CSR matrix into n_bands column bands.  band_rowPtr: [n_bands][rows + 1] (device, filled here);
// d_bcol / d_bval: device arrays of n_bands pointers to the per-band col / val arrays, which the caller allocates after
// band_nnz (host, filled by band_count) is known.  Two steps because the sizes come from the first.
void band_count(int rows, const int *d_rowPtr, const int *d_col, int band_cols, int n_bands, int *band_rowPtr,
                long long *band_nnz, cudaStream_t st) {
    if (n_bands > kMaxBands) throw std::runtime_error("too many column bands");
    const size_t stride = (size_t)rows + 1;
    band_count_kernel<<<(rows + 7) / 8, 256, 0, st>>>(d_rowPtr, d_col, rows, band_cols, n_bands, band_rowPtr);
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    HPR_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, band_rowPtr, band_rowPtr, (int)stride, st));
    HPR_CUDA_CHECK(cudaMallocAsync(&tmp, std::max<size_t>(tmp_bytes, 16), st));
    std::vector<int> last(n_bands);
    for (int b = 0; b < n_bands; ++b) {
        int *rp = band_rowPtr + (size_t)b * stride;     // counts[rows] is 0 (zero-filled by the caller) -> rp[rows] = total
        HPR_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, rp, rp, (int)stride, st));
        HPR_CUDA_CHECK(cudaMemcpyAsync(&last[b], rp + rows, sizeof(int), cudaMemcpyDeviceToHost, st));
    }
    HPR_CUDA_CHECK(cudaStreamSynchronize(st));
    cudaFreeAsync(tmp, st);
    for (int b = 0; b < n_bands; ++b) band_nnz[b] = last[b];
}
void band_fill(int rows, const int *d_rowPtr, const int *d_col, const double *d_val, int band_cols, int n_bands,
               const int *band_rowPtr, int *const *d_bcol, double *const *d_bval, cudaStream_t st) {
    band_fill_kernel<<<(rows + 7) / 8, 256, 0, st>>>(d_rowPtr, d_col, d_val, rows, band_cols, n_bands, band_rowPtr, d_bcol, d_bval);
    HPR_CUDA_CHECK(cudaGetLastError());
}

}  // namespace hpr
