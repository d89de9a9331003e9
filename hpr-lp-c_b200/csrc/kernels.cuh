// kernels.cuh -- hand-written fp64 sm_100a kernels of the HPR-LP engine.
//
// One streaming skeleton, `csr_stream_kernel`, carries every pass over a CSR matrix (the two
// fused HPR phases, the KKT residual passes, power iteration, Ruiz/Pock-Chambolle/Curtis-Reid
// row statistics).  It is an nnz-balanced ("merge-style") CSR-stream kernel:
//
//   item  = a fixed chunk of kChunk consecutive nonzeros (one CTA per item, grid = nnz/kChunk),
//           so the work per CTA is identical whatever the row-length distribution is
//           (power-law rows included; no row-id lists, no short/long buckets).
//   phase 1  every thread streams its nonzeros with 128-bit/64-bit coalesced loads
//           (double2 values, int2 column indices, evict-first), gathers the dense vector through
//           the read-only path (stays L2 resident) and stores the products in shared memory.
//   phase 2  G lanes per row (G chosen per matrix from the mean row length) sum the row's slice
//           of the product array; lane 0 runs the fused epilogue (projection, dual update,
//           Halpern averaging, residual terms ...).
//   rows cut by an item boundary publish a partial sum; the last contributor to arrive (one
//           atomic counter per row end) adds the partials in item order and runs the epilogue --
//           deterministic, and no second "fix-up" launch.
//
// Replaces, on the iteration path, the reference's fused_update_* kernels
// (src/cuda_kernels/HPR_cuda_kernels.cu:297-427), its cuSPARSE SpMV + elementwise kernels
// (src/main_iterate.cu:422-481, HPR_cuda_kernels.cu:203-295), the residual kernels + cuBLAS
// reductions (src/main_iterate.cu:207-309, HPR_cuda_kernels.cu:160-189) and advance_halpern_factors
// (HPR_cuda_kernels.cu:192-200).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace hpr {

constexpr int kThreads = 256;
constexpr int kPerThread = 8;
constexpr int kChunk = kThreads * kPerThread;  // nonzeros per item
constexpr int kMaxSlots = 8;                   // reduction slots per CTA

template <typename RP>
struct CsrView {
    int rows;
    long long nnz;
    const RP *rowPtr;
    const int *col;       // padded to a multiple of kChunk (pad: col 0, value 0)
    const double *val;    // padded likewise
    const int *item_row;  // n_items + 1 entries: first row finalised by item i
    int n_items;
    double *head_part;    // [n_items * 2] partial of the row entering the item from the left
    double *tail_part;    // [n_items * 2] partial of the row leaving the item to the right
    unsigned *counters;   // [n_items] arrivals per split row (indexed by the item where the row ends)
};

__device__ __forceinline__ double2 ld_stream(const double2 *p) { return __ldcs(p); }
__device__ __forceinline__ int2 ld_stream(const int2 *p) { return __ldcs(p); }

// Block-wide deterministic sum of NS per-thread accumulators; thread 0 writes them to
// out[blockIdx.x * kMaxSlots + s].  The caller's final_reduce_kernel adds the per-CTA values
// in block order, so every reduction is run-to-run reproducible.
template <int NS>
__device__ __forceinline__ void block_reduce_store(double (&acc)[NS], double *out, double *scratch /* >= 8*NS doubles */) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        double v = acc[s];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) scratch[s * (kThreads / 32) + wid] = v;
    }
    __syncthreads();
    if (threadIdx.x < NS) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) v += scratch[threadIdx.x * (kThreads / 32) + w];
        out[(size_t)blockIdx.x * kMaxSlots + threadIdx.x] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// The streaming skeleton.  Op supplies:
//   static constexpr int NV      number of product streams (1 or 2)
//   static constexpr bool kMax   combine with fmax instead of +
//   void init()                  per-thread scalar loads
//   void elem(v, col, out[NV])   per-nonzero term
//   void row(r, acc[NV], p0, p1) fused epilogue of a complete row
//   void finish(scratch)         CTA-level reductions (may be empty)
// ------------------------------------------------------------------------------------------------
template <class Op, int G, typename RP>
__global__ void __launch_bounds__(kThreads, 4) csr_stream_kernel(CsrView<RP> M, Op op) {
    constexpr int NV = Op::NV;
    __shared__ __align__(16) double prod[NV][kChunk];
    __shared__ double red_scratch[kMaxSlots * (kThreads / 32)];

    op.init();
    const long long s = (long long)blockIdx.x * kChunk;
    const long long e = (s + kChunk < M.nnz) ? s + kChunk : M.nnz;

    // ---- phase 1: stream nonzeros, gather, multiply ------------------------------------------------
    {
        const double2 *v2 = reinterpret_cast<const double2 *>(M.val + s);
        const int2 *c2 = reinterpret_cast<const int2 *>(M.col + s);
        double2 vv[kPerThread / 2];
        int2 cc[kPerThread / 2];
#pragma unroll
        for (int u = 0; u < kPerThread / 2; ++u) {
            cc[u] = ld_stream(c2 + u * kThreads + threadIdx.x);
            vv[u] = ld_stream(v2 + u * kThreads + threadIdx.x);
        }
#pragma unroll
        for (int u = 0; u < kPerThread / 2; ++u) {
            double o0[NV], o1[NV];
            op.elem(vv[u].x, cc[u].x, o0);
            op.elem(vv[u].y, cc[u].y, o1);
            const int t = u * kThreads + threadIdx.x;
#pragma unroll
            for (int q = 0; q < NV; ++q)
                *reinterpret_cast<double2 *>(&prod[q][2 * t]) = make_double2(o0[q], o1[q]);
        }
    }
    __syncthreads();

    // ---- phase 2: per-row sums + fused epilogue ------------------------------------------------------
    const int rA = M.item_row[blockIdx.x];
    const int rB = M.item_row[blockIdx.x + 1];
    const int lane = threadIdx.x & 31;
    const int gl = threadIdx.x & (G - 1);
    const int gid = threadIdx.x / G;
    constexpr int NGRP = kThreads / G;
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(G - 1)));

    for (int base = rA; base <= rB; base += NGRP) {
        const int r = base + gid;
        const bool valid = (r <= rB) && (r < M.rows);
        long long p0 = 0, p1 = 0;
        int lo = 0, hi = 0;
        if (valid) {
            p0 = (long long)M.rowPtr[r];
            p1 = (long long)M.rowPtr[r + 1];
            const long long a = p0 > s ? p0 : s;
            const long long b = p1 < e ? p1 : e;
            if (b > a) { lo = (int)(a - s); hi = (int)(b - s); }
        }
        double acc[NV];
#pragma unroll
        for (int q = 0; q < NV; ++q) acc[q] = 0.0;   // |a| >= 0, so 0 is also the identity of fmax here
        for (int k = lo + gl; k < hi; k += G) {
#pragma unroll
            for (int q = 0; q < NV; ++q) acc[q] = Op::kMax ? fmax(acc[q], prod[q][k]) : acc[q] + prod[q][k];
        }
#pragma unroll
        for (int off = G / 2; off > 0; off >>= 1) {
#pragma unroll
            for (int q = 0; q < NV; ++q) {
                const double o = __shfl_xor_sync(gmask, acc[q], off);
                acc[q] = Op::kMax ? fmax(acc[q], o) : acc[q] + o;
            }
        }
        if (!valid) continue;   // whole group leaves together (r is group-uniform)

        const bool head = (r == rA) && (p0 < s);   // row entered this item from the left
        const bool cont = (p1 > e);                // row continues to the right
        if (!head && !cont) {
            if (gl == 0) op.row(r, acc, p0, p1);
            continue;
        }
        if (!head && p0 >= e) continue;            // r == rB but it starts in a later item
        // ---- split row: publish the partial, last arriver finalises -----------------------------
        const int ia = (int)(p0 / kChunk);
        const int ib = (int)((p1 - 1) / kChunk);
        int last = 0;
        if (gl == 0) {
            double *slot = (head ? M.head_part : M.tail_part) + (size_t)blockIdx.x * 2;
#pragma unroll
            for (int q = 0; q < NV; ++q) slot[q] = acc[q];
            __threadfence();
            const unsigned old = atomicAdd(&M.counters[ib], 1u);
            last = (old == (unsigned)(ib - ia));
        }
        last = __shfl_sync(gmask, last, lane & ~(G - 1));
        if (last) {
            __threadfence();
            double tot[NV];
#pragma unroll
            for (int q = 0; q < NV; ++q) tot[q] = 0.0;
            for (int k = ia + gl; k <= ib; k += G) {
                const double *src = ((k == ia) ? M.tail_part : M.head_part) + (size_t)k * 2;
#pragma unroll
                for (int q = 0; q < NV; ++q) {
                    const double pv = __ldcg(src + q);
                    tot[q] = Op::kMax ? fmax(tot[q], pv) : tot[q] + pv;
                }
            }
#pragma unroll
            for (int off = G / 2; off > 0; off >>= 1) {
#pragma unroll
                for (int q = 0; q < NV; ++q) {
                    const double o = __shfl_xor_sync(gmask, tot[q], off);
                    tot[q] = Op::kMax ? fmax(tot[q], o) : tot[q] + o;
                }
            }
            if (gl == 0) {
                M.counters[ib] = 0u;   // re-arm for the next launch
                op.row(r, tot, p0, p1);
            }
        }
    }
    op.finish(red_scratch);
}

// first row finalised by item i = first r with rowPtr[r+1] > i*kChunk (item 0 also owns leading
// empty rows, item n_items-1 the trailing ones).
template <typename RP>
__global__ void build_item_rows_kernel(const RP *rowPtr, int rows, int n_items, int *item_row) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n_items) return;
    if (i == 0) { item_row[0] = 0; return; }
    if (i == n_items) { item_row[i] = rows; return; }
    const long long target = (long long)i * kChunk;
    int lo = 0, hi = rows;
    while (lo < hi) {
        const int mid = lo + ((hi - lo) >> 1);
        if ((long long)rowPtr[mid + 1] > target) hi = mid; else lo = mid + 1;
    }
    item_row[i] = lo;
}

// Sum the per-CTA partials (block order) into out[0..ns): one CTA, fixed tree => deterministic.
__global__ void final_reduce_kernel(const double *partials, int n_blocks, int ns, double *out);

// ================================================================================================
// Ops
// ================================================================================================
struct OpBase {
    static constexpr int NV = 1;
    static constexpr bool kMax = false;
    __device__ __forceinline__ void init() {}
    __device__ __forceinline__ void finish(double *) {}
};

// x-phase (reference fused_update_x_z_rows_*_kernel, HPR_cuda_kernels.cu:297-361; check variant
// update_zx_check_kernel :203-226):  w = (A^T y)_j ; zt = x + sigma (w - c) ; x_bar = proj_[l,u] zt ;
// x_hat = 2 x_bar - x ; x <- f2 x_hat + f1 x0 ; check also stores x_bar, z_bar=(x_bar-zt)/sigma, x_bar-x_hat.
// Halpern counter: this kernel reads k from kx and mirrors it into ky for the y-phase.
template <bool CHECK>
struct XPhaseOp : OpBase {
    const double *y;
    double *x, *x_hat;
    const double *c, *l, *u, *x0;
    double *x_bar, *z_bar, *x_tmp;
    const double *params;   // [sigma, lambda*sigma, 1/(lambda*sigma), 1/sigma]
    const int *kx;
    int *ky;
    double sigma, f1, f2;
    __device__ __forceinline__ void init() {
        sigma = params[0];
        const int k = *kx;
        f1 = 1.0 / (k + 2.0);
        f2 = 1.0 - f1;
        if (blockIdx.x == 0 && threadIdx.x == 0) *ky = k;
    }
    __device__ __forceinline__ void elem(double v, int col, double (&o)[1]) const { o[0] = v * __ldg(y + col); }
    __device__ __forceinline__ void row(int j, const double (&acc)[1], long long, long long) const {
        const double xi = x[j];
        const double zt = fma(sigma, acc[0] - c[j], xi);
        const double xb = fmin(u[j], fmax(l[j], zt));
        const double xh = 2.0 * xb - xi;
        x[j] = fma(f2, xh, f1 * x0[j]);
        x_hat[j] = xh;
        if (CHECK) {
            x_bar[j] = xb;
            z_bar[j] = (xb - zt) / sigma;
            x_tmp[j] = xb - xh;
        }
    }
};

// y-phase (reference fused_update_y_rows_*_kernel :363-427; check variant update_y_check_kernel :249-272):
// v = (A x_hat)_i - lambda sigma y ; d = max(AL - v, min(AU - v, 0)) ; y_bar = d/(lambda sigma) ;
// y_hat = 2 y_bar - y ; y <- f2 y_hat + f1 y0 ; check also stores y_bar, y_obj = v + d, y_bar - y_hat.
// Advances the Halpern counter: kx <- ky + 1 (reference advance_halpern_factors_kernel :192-200).
template <bool CHECK>
struct YPhaseOp : OpBase {
    const double *x_hat;
    double *y;
    const double *AL, *AU, *y0;
    double *y_bar, *y_obj, *y_tmp;
    const double *params;
    const int *ky;
    int *kx;
    double lamsig, inv_lamsig, f1, f2;
    __device__ __forceinline__ void init() {
        lamsig = params[1];
        inv_lamsig = params[2];
        const int k = *ky;
        f1 = 1.0 / (k + 2.0);
        f2 = 1.0 - f1;
        if (blockIdx.x == 0 && threadIdx.x == 0) *kx = k + 1;
    }
    __device__ __forceinline__ void elem(double v, int col, double (&o)[1]) const { o[0] = v * __ldg(x_hat + col); }
    __device__ __forceinline__ void row(int i, const double (&acc)[1], long long, long long) const {
        const double yi = y[i];
        const double v = fma(-lamsig, yi, acc[0]);
        const double d = fmax(AL[i] - v, fmin(AU[i] - v, 0.0));
        const double yb = inv_lamsig * d;
        const double yh = 2.0 * yb - yi;
        y[i] = fma(f2, yh, f1 * y0[i]);
        if (CHECK) {
            y_bar[i] = yb;
            y_obj[i] = v + d;
            y_tmp[i] = yb - yh;
        }
    }
};

// Dual residual pass over A^T (reference residual_compute_Rd_cusparse + queue_dot/nrm2,
// src/main_iterate.cu:218-226,237-258): slots 0 |Rd|^2, 1 <c,x_bar>, 2 <x_bar,z_bar>, 3 |x_tmp|^2 (gap),
// 4 |bound violation / col_norm|^2 (iteration 0, reference residual_compute_lu_kernel :174-180).
template <bool GAP, bool ITER0>
struct ResidualDualOp : OpBase {
    const double *y_bar, *c, *z_bar, *x_bar, *x_tmp, *col_norm, *l, *u;
    double *partials;
    double t[5];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int s = 0; s < 5; ++s) t[s] = 0.0;
    }
    __device__ __forceinline__ void elem(double v, int col, double (&o)[1]) const { o[0] = v * __ldg(y_bar + col); }
    __device__ __forceinline__ void row(int j, const double (&acc)[1], long long, long long) {
        const double cj = c[j], zb = z_bar[j], xb = x_bar[j], cn = col_norm[j];
        const double rd = (cj - acc[0] - zb) * cn;
        t[0] += rd * rd;
        t[1] += cj * xb;
        t[2] += xb * zb;
        if (GAP) { const double dx = x_tmp[j]; t[3] += dx * dx; }
        if (ITER0) {
            const double lj = l[j], uj = u[j];
            const double viol = (xb < lj) ? (lj - xb) : ((xb > uj) ? (xb - uj) : 0.0);
            const double q = viol / cn;
            t[4] += q * q;
        }
    }
    __device__ __forceinline__ void finish(double *scratch) { block_reduce_store<5>(t, partials, scratch); }
};

// Primal residual pass over A (reference residual_compute_Rp_cusparse, src/main_iterate.cu:207-215,
// plus the restart-gap SpMV/dots :245-254): slots 0 |Rp|^2, 1 <y_obj,y_bar>, 2 <A x_tmp, y_tmp>, 3 |y_tmp|^2.
template <bool GAP>
struct ResidualPrimalOp : OpBase {
    static constexpr int NV = GAP ? 2 : 1;
    const double *x_bar, *x_tmp, *AL, *AU, *row_norm, *y_obj, *y_bar, *y_tmp;
    double *partials;
    double t[4];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int s = 0; s < 4; ++s) t[s] = 0.0;
    }
    __device__ __forceinline__ void elem(double v, int col, double (&o)[NV]) const {
        o[0] = v * __ldg(x_bar + col);
        if (GAP) o[NV - 1] = v * __ldg(x_tmp + col);
    }
    __device__ __forceinline__ void row(int i, const double (&acc)[NV], long long, long long) {
        const double ax = acc[0];
        const double rp = fmax(fmin(AU[i] - ax, 0.0), AL[i] - ax) * row_norm[i];
        t[0] += rp * rp;
        t[1] += y_obj[i] * y_bar[i];
        if (GAP) { const double dy = y_tmp[i]; t[2] += acc[NV - 1] * dy; t[3] += dy * dy; }
    }
    __device__ __forceinline__ void finish(double *scratch) { block_reduce_store<4>(t, partials, scratch); }
};

// M-norm cross term after a restart iteration (reference compute_weighted_norm,
// src/main_iterate.cu:486-515): slots 0 <A dx, dy>, 1 |dy|^2.
struct WeightedNormOp : OpBase {
    const double *dx, *dy;
    double *partials;
    double t[2];
    __device__ __forceinline__ void init() { t[0] = t[1] = 0.0; }
    __device__ __forceinline__ void elem(double v, int col, double (&o)[1]) const { o[0] = v * __ldg(dx + col); }
    __device__ __forceinline__ void row(int i, const double (&acc)[1], long long, long long) {
        const double d = dy[i];
        t[0] += acc[0] * d;
        t[1] += d * d;
    }
    __device__ __forceinline__ void finish(double *scratch) { block_reduce_store<2>(t, partials, scratch); }
};

// Plain SpMV out = M * g, with optional fused <out,out> and <q,out> (power iteration,
// reference src/power_iteration.cu:73-90).
template <bool DOTS>
struct SpmvOp : OpBase {
    const double *g;
    double *out;
    const double *q;
    double *partials;
    double t[2];
    __device__ __forceinline__ void init() { t[0] = t[1] = 0.0; }
    __device__ __forceinline__ void elem(double v, int col, double (&o)[1]) const { o[0] = v * __ldg(g + col); }
    __device__ __forceinline__ void row(int i, const double (&acc)[1], long long, long long) {
        out[i] = acc[0];
        if (DOTS) { t[0] += acc[0] * acc[0]; t[1] += q[i] * acc[0]; }
    }
    __device__ __forceinline__ void finish(double *scratch) {
        if (DOTS) block_reduce_store<2>(t, partials, scratch);
    }
};

// Row statistic for Ruiz (sqrt max|a|) and Pock-Chambolle (sqrt sum|a|) scaling
// (reference CSR_A_row_norm_kernel, HPR_cuda_kernels.cu:91-120); <1e-15 -> 1.
template <bool MAX>
struct RowNormOp : OpBase {
    static constexpr bool kMax = MAX;
    double *out;
    __device__ __forceinline__ void elem(double v, int, double (&o)[1]) const { o[0] = fabs(v); }
    __device__ __forceinline__ void row(int i, const double (&acc)[1], long long, long long) const {
        double r = sqrt(acc[0]);
        if (r < 1e-15) r = 1.0;
        out[i] = r;
    }
};

// Curtis-Reid log-domain sweep (reference curtis_reid_log_update_kernel, src/scaling.cu:5-31).
struct CurtisReidOp : OpBase {
    const double *other;
    double *out;
    __device__ __forceinline__ void elem(double v, int col, double (&o)[1]) const {
        o[0] = -log(fmax(fabs(v), 1e-300)) - __ldg(other + col);
    }
    __device__ __forceinline__ void row(int i, const double (&acc)[1], long long p0, long long p1) const {
        const long long cnt = p1 - p0;
        out[i] = cnt > 0 ? acc[0] / (double)cnt : 0.0;
    }
};

// ------------------------------------------------------------------------------------------------
// Matrix value scaling, one pass: v <- (v op f_first) op f_second with two separately rounded
// operations in the reference's order (mul_CSR_A_row then mul_CSR_AT_row, HPR_cuda_kernels.cu:122-157;
// call order src/scaling.cu:72-76,136-141): for A   first = rowfac[row],  second = gathfac[col];
//                                           for A^T first = gathfac[col], second = rowfac[row].
// Same item decomposition as csr_stream_kernel: row factors are expanded into shared memory by the
// row owners, then the nnz-parallel phase is fully coalesced.
// ------------------------------------------------------------------------------------------------
template <bool DIVIDE, bool ROW_FIRST, typename RP>
__global__ void __launch_bounds__(kThreads, 4)
scale_values_kernel(CsrView<RP> M, double *val_rw, const double *rowfac, const double *gathfac) {
    __shared__ double rf[kChunk];
    const long long s = (long long)blockIdx.x * kChunk;
    const long long e = (s + kChunk < M.nnz) ? s + kChunk : M.nnz;
    const int rA = M.item_row[blockIdx.x];
    const int rB = M.item_row[blockIdx.x + 1];
    constexpr int GG = 8;
    const int gl = threadIdx.x & (GG - 1);
    const int gid = threadIdx.x / GG;
    for (int r = rA + gid; r <= rB && r < M.rows; r += kThreads / GG) {
        const long long p0 = (long long)M.rowPtr[r], p1 = (long long)M.rowPtr[r + 1];
        const long long a = p0 > s ? p0 : s, b = p1 < e ? p1 : e;
        const double f = rowfac[r];
        for (long long k = a + gl; k < b; k += GG) rf[(int)(k - s)] = f;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < (int)(e - s); t += kThreads) {
        double v = val_rw[s + t];
        const double fr = rf[t];
        const double fg = __ldg(gathfac + M.col[s + t]);
        const double f1 = ROW_FIRST ? fr : fg;
        const double f2 = ROW_FIRST ? fg : fr;
        if (DIVIDE) { v = v / f1; v = v / f2; } else { v = v * f1; v = v * f2; }
        val_rw[s + t] = v;
    }
}

}  // namespace hpr
