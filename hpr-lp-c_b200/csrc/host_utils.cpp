// host_utils.cpp -- multi-threaded host helpers of the model layer (create_model_from_arrays): the CSC -> CSR conversion
// the Julia and MATLAB bindings go through (they hold A column-major and pass is_csc = true; reference
// src/HPRLP.cu:354-396 does it with a single-threaded counting sort) and the copies of the caller's arrays.
#include <algorithm>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "engine.h"

namespace hpr {

static int host_thread_count(long long work, long long per_thread) {
    int t = 1;
#ifdef _OPENMP
    t = std::max(1, omp_get_max_threads());
#endif
    return (int)std::max<long long>(1, std::min<long long>(t, work / std::max<long long>(per_thread, 1)));
}

// dst = src, split over the host threads for large arrays
void copy_mt(void *dst, const void *src, size_t bytes) {
    const int T = host_thread_count((long long)bytes, 8 << 20);
    if (T <= 1) { std::memcpy(dst, src, bytes); return; }
#pragma omp parallel for schedule(static, 1) num_threads(T)
    for (int t = 0; t < T; ++t) {
        const size_t lo = bytes * (size_t)t / T, hi = bytes * (size_t)(t + 1) / T;
        std::memcpy(static_cast<char *>(dst) + lo, static_cast<const char *>(src) + lo, hi - lo);
    }
}

// Same result, entry for entry, as csr_transpose_host (stable counting sort by column: the entries of a transposed row
// keep the source row order), with the source rows split into nnz-balanced slices: slice t counts its entries per
// column, a prefix over (column, slice) gives every slice its first slot in every transposed row, each slice then
// scatters its own entries in order.
void csr_transpose_host_mt(int rows, int cols, int nnz, const int *rp, const int *ci, const double *v, int *trp, int *tci,
                           double *tv) {
    int T = host_thread_count(nnz, 1 << 20);
    while (T > 1 && (size_t)T * (size_t)cols * sizeof(int) > ((size_t)1 << 30)) T /= 2;   // counters: T x cols ints
    if (T <= 1) { csr_transpose_host(rows, cols, nnz, rp, ci, v, trp, tci, tv); return; }
    std::vector<int> rb(T + 1, 0);   // slice t = source rows [rb[t], rb[t+1])
    rb[T] = rows;
    for (int t = 1; t < T; ++t) {
        const int target = (int)((long long)nnz * t / T);
        rb[t] = (int)(std::lower_bound(rp, rp + rows + 1, target) - rp);
        rb[t] = std::min(std::max(rb[t], rb[t - 1]), rows);
    }
    std::vector<std::vector<int>> cnt(T);
#pragma omp parallel for schedule(static, 1) num_threads(T)
    for (int t = 0; t < T; ++t) {
        cnt[t].assign((size_t)cols, 0);
        for (int k = rp[rb[t]]; k < rp[rb[t + 1]]; ++k) cnt[t][ci[k]]++;
    }
    trp[0] = 0;
    for (int c = 0; c < cols; ++c) {
        int run = trp[c];
        for (int t = 0; t < T; ++t) { const int x = cnt[t][c]; cnt[t][c] = run; run += x; }
        trp[c + 1] = run;
    }
#pragma omp parallel for schedule(static, 1) num_threads(T)
    for (int t = 0; t < T; ++t) {
        int *cur = cnt[t].data();
        for (int i = rb[t]; i < rb[t + 1]; ++i)
            for (int k = rp[i]; k < rp[i + 1]; ++k) {
                const int pos = cur[ci[k]]++;
                tci[pos] = i;
                tv[pos] = v[k];
            }
    }
}

}  // namespace hpr
