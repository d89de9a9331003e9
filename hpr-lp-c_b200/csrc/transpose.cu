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

}  // namespace

void device_transpose_csr(int rows, int cols, int nnz, const int *d_rowPtr, const int *d_col, const double *d_val,
                          int *d_trp, int *d_tcol, double *d_tval, cudaStream_t st) {
    int *rowid = nullptr, *idx = nullptr, *keys_out = nullptr, *perm = nullptr;
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    // stream-ordered temporaries from the device memory pool (cached across solves, no synchronous cudaMalloc/cudaFree)
    HPR_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void **>(&rowid), sizeof(int) * (size_t)nnz, st));
    HPR_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void **>(&idx), sizeof(int) * (size_t)nnz, st));
    HPR_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void **>(&keys_out), sizeof(int) * (size_t)nnz, st));
    HPR_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void **>(&perm), sizeof(int) * (size_t)nnz, st));
    const int T = 256;
    nnz_row_ids_kernel<<<(nnz + T - 1) / T, T, 0, st>>>(d_rowPtr, rows, nnz, rowid, idx);
    int bits = 1;
    while (bits < 31 && (1LL << bits) < (long long)cols) ++bits;
    HPR_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_col, keys_out, idx, perm, nnz, 0, bits, st));
    HPR_CUDA_CHECK(cudaMallocAsync(&tmp, std::max<size_t>(tmp_bytes, 16), st));
    HPR_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, d_col, keys_out, idx, perm, nnz, 0, bits, st));
    permute_kernel<<<(nnz + T - 1) / T, T, 0, st>>>(perm, rowid, d_val, nnz, d_tcol, d_tval);
    col_ptr_kernel<<<(cols + 1 + T - 1) / T, T, 0, st>>>(keys_out, nnz, cols, d_trp);
    cudaFreeAsync(rowid, st); cudaFreeAsync(idx, st); cudaFreeAsync(keys_out, st); cudaFreeAsync(perm, st); cudaFreeAsync(tmp, st);
}

}  // namespace hpr
