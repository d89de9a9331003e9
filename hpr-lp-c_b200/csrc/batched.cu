// batched.cu -- solve_batched / free_batched_results: B LPs sharing one sparse matrix A
// (reference src/batched_solver.cu:939-1105 and its kernels :122-323).
//
// B200-native design (replaces cuSPARSE SpMM on column-major n x B matrices + flat elementwise kernels +
// per-instance blocking cuBLAS reductions + a host sync every iteration):
//   * device layout [group][row][32]: 32 instances ("a group") are contiguous, so every access to a
//     batched vector is a 256-byte fully coalesced warp transaction; lane = instance.
//   * fused SpMM + prox kernels: one warp per matrix row; the row's (col,val) pairs are loaded
//     cooperatively once and broadcast by shuffle, each gather is one coalesced 256 B read of the
//     dense row, the projection / dual update / Halpern averaging run in the epilogue -- A^T Y and
//     A X_hat are never written to memory.  A is streamed once per group and stays L2 resident
//     across groups (C4: 24 MB per copy vs 126 MB L2); groups are scheduled group-major so the
//     gathered slab of one group (n x 32 doubles) is the L2 working set.
//   * per-instance reductions (KKT residuals, objectives, restart gaps, movement norms) are
//     accumulated per lane inside the same passes and reduced in a fixed order: one D2H of
//     slots x B doubles per check instead of 5-6 blocking cuBLAS calls per instance.
//   * per-instance sigma, Halpern counters and active/restart masks live on the device; the host
//     only runs the restart/sigma/stopping logic at check points (no per-iteration sync).
// Host logic (per-instance scaling with long-double norms, +-inf -> +-1e100, restart rules, sigma
// update, `<=` stopping test, shared lambda_max bumped by max) restates the reference line by line.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../include/batched_solver.h"
#include "abi_guard.h"
#include "engine.h"
#include "kernels.cuh"

namespace hpr {
namespace {

constexpr int kGS = 32;            // instances per group = lanes per warp
constexpr int kBThreads = 256;     // 8 warps per CTA
constexpr int kBWarps = kBThreads / 32;
#ifndef HPR_B_ROWS
#define HPR_B_ROWS 32
#endif
#ifndef HPR_B_MINB
#define HPR_B_MINB 6
#endif
#ifndef HPR_B_PIPE
#define HPR_B_PIPE 0
#endif
#ifndef HPR_B_UNROLL
#define HPR_B_UNROLL 4
#endif
#ifndef HPR_B_ROWS2
#define HPR_B_ROWS2 0   // 1: a warp works on two rows at once (twice the independent loads in flight per warp)
#endif
#define HPR_PRAGMA_(x) _Pragma(#x)
#define HPR_UNROLL_(n) HPR_PRAGMA_(unroll n)
constexpr int kRowsPerCta = HPR_B_ROWS;    // rows per CTA (8 warps): the software pipeline of the row loop needs a few rows per warp to fill
constexpr double kInfReplacement = 1.0e100;   // reference src/batched_solver.cu:17

#ifndef HPR_B_SMEM_BCAST
#define HPR_B_SMEM_BCAST 1
#endif
struct alignas(16) BPair { double v; int c; int pad; };

struct BView {
    int rows;          // rows of this matrix
    int gcols;         // rows of the gathered dense operand (= columns of this matrix)
    const int *rowPtr;
    const int *col;
    const double *val;
};

// L2 residency policy of the batched passes (north_star: "gathered vectors staged through ... L2-persisting windows").
// The gathered operand of a pass (one [row][32] slab per group: 12.8 MB for Y, 51 MB for X_hat on configs[3]) is re-read
// ~nnz/rows times and must stay in the 126 MB L2; everything else is touched once per pass.  So gathers are loaded with
// evict_last priority and the per-row vectors with evict_first / streaming stores.  Without the hints the streamed
// vectors evict the slab: ncu on r1 showed a 47 % L2 hit rate and 3.4x the algorithmic DRAM reads in the y-phase.
__device__ __forceinline__ unsigned long long make_keep_policy() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ double ld_keep(const double *p, unsigned long long pol) {
    double v;
    asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ double ld_once(const double *p) { return __ldcs(p); }
__device__ __forceinline__ void st_once(double *p, double v) { __stcs(p, v); }

// ---------------------------------------------------------------------------------------------
// skeleton: one warp per row, lane = instance of group blockIdx.y
// ---------------------------------------------------------------------------------------------
// Register blocking over the batch: one pass over the matrix can serve Op::kNG groups of 32 instances (lane = instance
// inside each group), so every (col, val) pair -- loaded once per 32 nonzeros and broadcast by two shuffles -- feeds kNG
// independent gathers + FMAs per lane and A is streamed once per kNG groups.  Built and measured in round 2: it does NOT
// pay on B200 (see BatchedSolver::ngx) -- the passes are limited by the bytes the gathers pull through the L2, which
// register blocking does not reduce, and kNG slabs thrash the L2 -- so kNG = 1 is the default.
// The row loop is software-pipelined (r2): a row costs a CHAIN of memory round trips -- row extent -> (col, val) ->
// gathers -> epilogue operands -- and ncu showed the r1 kernel waiting on that chain with nothing saturated (L2 34-39 %,
// L1TEX 49-59 %, DRAM 32-46 % of peak).  Here the epilogue operands of the row (Op::pre) and the NEXT row's extent are
// requested before the gathers, and the next row's first 32 (col, val) pairs under them: one exposed round trip per row
// (the gathers) instead of four.
template <class Op>
__global__ void __launch_bounds__(kBThreads, Op::kNG >= 4 ? 3 : (Op::kNG == 2 ? 4 : HPR_B_MINB)) batched_rows_kernel(BView M, Op op, int G) {
    constexpr int NG = Op::kNG;
    __shared__ double red[kBWarps][kMaxSlots][kGS];
#if HPR_B_SMEM_BCAST
    __shared__ BPair pairs[kBWarps][32];
#endif
    const int g0 = blockIdx.y * NG;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    op.init(g0, lane, G);
    const unsigned long long keep = make_keep_policy();
    const int row0 = blockIdx.x * kRowsPerCta;
    const int row1 = min(M.rows, row0 + kRowsPerCta);
#if HPR_B_ROWS2
    if constexpr (NG == 1) {
        for (int rA = row0 + warp; rA < row1; rA += 2 * kBWarps) {
            const int rB = rA + kBWarps;
            const bool hasB = rB < row1;
            const int a0 = M.rowPtr[rA], a1 = M.rowPtr[rA + 1];
            const int b0 = hasB ? M.rowPtr[rB] : 0, b1 = hasB ? M.rowPtr[rB + 1] : 0;
            double accA = 0.0, accB = 0.0;
            const int lenA = a1 - a0, lenB = b1 - b0;
            for (int k0 = 0; k0 < lenA || k0 < lenB; k0 += 32) {
                const int ka = a0 + k0 + lane, kb = b0 + k0 + lane;
                const int cA = (ka < a1) ? __ldg(M.col + ka) : 0, cB = (kb < b1) ? __ldg(M.col + kb) : 0;
                const double vA = (ka < a1) ? __ldg(M.val + ka) : 0.0, vB = (kb < b1) ? __ldg(M.val + kb) : 0.0;
                const int cntA = max(0, min(32, lenA - k0)), cntB = max(0, min(32, lenB - k0));
                const int cnt = max(cntA, cntB);
HPR_UNROLL_(HPR_B_UNROLL)
                for (int t = 0; t < cnt; ++t) {
                    const int ccA = __shfl_sync(0xffffffffu, cA, t), ccB = __shfl_sync(0xffffffffu, cB, t);
                    const double vvA = __shfl_sync(0xffffffffu, vA, t), vvB = __shfl_sync(0xffffffffu, vB, t);
                    if (t < cntA) op.accum(vvA, ccA, 0, accA, keep);
                    if (t < cntB) op.accum(vvB, ccB, 0, accB, keep);
                }
            }
            const typename Op::Pre preA = op.pre(rA, 0);
            typename Op::Pre preB{};
            if (hasB) preB = op.pre(rB, 0);
            op.row(rA, 0, accA, preA);
            if (hasB) op.row(rB, 0, accB, preB);
        }
        op.finish(red, warp, lane);
        return;
    }
#endif
    int r = row0 + warp;
    int p0 = 0, p1 = 0, c = 0;
    double v = 0.0;
    if (r < row1) {
        p0 = M.rowPtr[r]; p1 = M.rowPtr[r + 1];
        if (p0 + lane < p1) { c = __ldg(M.col + p0 + lane); v = __ldg(M.val + p0 + lane); }
    }
    while (r < row1) {
        const int rn = r + kBWarps;
        int np0 = 0, np1 = 0, nc = 0;
        double nv = 0.0;
#if HPR_B_PIPE
        if (rn < row1) { np0 = M.rowPtr[rn]; np1 = M.rowPtr[rn + 1]; }   // next row's extent: in flight under this row's gathers
#endif
#if HPR_B_PIPE == 1
        typename Op::Pre pre[NG];
#pragma unroll
        for (int q = 0; q < NG; ++q) pre[q] = op.pre(r, q);              // epilogue operands of this row: likewise
#endif
        double acc[NG];
#pragma unroll
        for (int q = 0; q < NG; ++q) acc[q] = 0.0;
        for (int k0 = p0; k0 < p1; k0 += 32) {
            if (k0 != p0) {                                               // rows longer than 32 nonzeros: the later chunks
                const int kk = k0 + lane;
                c = (kk < p1) ? __ldg(M.col + kk) : 0;
                v = (kk < p1) ? __ldg(M.val + kk) : 0.0;
            }
            const int cnt = min(32, p1 - k0);
#if HPR_B_SMEM_BCAST
            // the 32 (col, val) pairs of the block go through shared memory: one 16-byte broadcast read per nonzero instead
            // of three shuffles (a 64-bit value is two) on the L1 data stage
            __syncwarp();
            pairs[warp][lane] = BPair{v, c, 0};
            __syncwarp();
HPR_UNROLL_(HPR_B_UNROLL)
            for (int t = 0; t < cnt; ++t) {
                const BPair pr = pairs[warp][t];
#pragma unroll
                for (int q = 0; q < NG; ++q) op.accum(pr.v, pr.c, q, acc[q], keep);
            }
#else
HPR_UNROLL_(HPR_B_UNROLL)
            for (int t = 0; t < cnt; ++t) {
                const int cc = __shfl_sync(0xffffffffu, c, t);
                const double vv = __shfl_sync(0xffffffffu, v, t);
#pragma unroll
                for (int q = 0; q < NG; ++q) op.accum(vv, cc, q, acc[q], keep);
            }
#endif
#if HPR_B_PIPE == 1
            if (k0 == p0 && np0 + lane < np1) { nc = __ldg(M.col + np0 + lane); nv = __ldg(M.val + np0 + lane); }
#endif
        }
#if HPR_B_PIPE == 1
        if (p1 <= p0 && np0 + lane < np1) { nc = __ldg(M.col + np0 + lane); nv = __ldg(M.val + np0 + lane); }   // (empty row)
#else
#if HPR_B_PIPE == 0
        if (rn < row1) { np0 = M.rowPtr[rn]; np1 = M.rowPtr[rn + 1]; }
#endif
        // after the gathers have been issued: the next row's first 32 (col, val) pairs and this row's epilogue operands
        if (np0 + lane < np1) { nc = __ldg(M.col + np0 + lane); nv = __ldg(M.val + np0 + lane); }
        typename Op::Pre pre[NG];
#pragma unroll
        for (int q = 0; q < NG; ++q) pre[q] = op.pre(r, q);
#endif
#pragma unroll
        for (int q = 0; q < NG; ++q) op.row(r, q, acc[q], pre[q]);
        r = rn; p0 = np0; p1 = np1; c = nc; v = nv;
    }
    op.finish(red, warp, lane);
}

template <int NS>
__device__ __forceinline__ void batched_reduce_store(const double (&t)[NS], double (*red)[kMaxSlots][kGS], int warp, int lane,
                                                     double *partials) {
#pragma unroll
    for (int s = 0; s < NS; ++s) red[warp][s][lane] = t[s];
    __syncthreads();
    if (warp == 0) {
        const size_t blk = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            double v = 0.0;
#pragma unroll
            for (int w = 0; w < kBWarps; ++w) v += red[w][s][lane];
            partials[(blk * kMaxSlots + s) * kGS + lane] = v;
        }
    }
}

// out[s * Bpad + g*32 + lane] = sum over the nbx CTAs of group g, in block order
__global__ void batched_final_reduce_kernel(const double *partials, int nbx, int ns, int Bpad, double *out) {
    const int g = blockIdx.x, lane = threadIdx.x & 31, s = threadIdx.x >> 5;
    if (s >= ns) return;
    double v = 0.0;
    for (int b = 0; b < nbx; ++b) v += partials[(((size_t)g * nbx + b) * kMaxSlots + s) * kGS + lane];
    out[(size_t)s * Bpad + g * kGS + lane] = v;
}

// Per-thread state shared by every op: where the gathered operand and the row-indexed arrays of each of the NG groups
// start for this lane.  Groups beyond G (last pass of a batch whose group count is not a multiple of NG) alias group g0
// for the gathers and are switched off for the epilogue.
template <int NG>
struct BOpBase {
    static constexpr int kNG = NG;
    int gcols, orows;               // rows of the gathered operand / of the output-side arrays
    const double *gp[NG];           // gathered operand of group q at row 0, this lane
    size_t obase[NG];               // offset of (group q, row 0, this lane) in the row-indexed arrays
    bool valid[NG];
    __device__ __forceinline__ void init_base(const double *gathered, int g0, int lane, int G) {
#pragma unroll
        for (int q = 0; q < NG; ++q) {
            valid[q] = g0 + q < G;
            const int g = valid[q] ? g0 + q : g0;
            gp[q] = gathered + (size_t)g * gcols * kGS + lane;
            obase[q] = (size_t)g * orows * kGS + lane;
        }
    }
    __device__ __forceinline__ void accum(double v, int col, int q, double &acc, unsigned long long keep) const {
        acc = fma(v, ld_keep(gp[q] + (size_t)col * kGS, keep), acc);
    }
    __device__ __forceinline__ void finish(double (*)[kMaxSlots][kGS], int, int) {}
    struct Pre {};   // epilogue operands requested before the gathers (ops that have none)
    __device__ __forceinline__ Pre pre(int, int) const { return Pre{}; }
};

// x-phase (reference update_x_z_{normal,check}_batched_kernel, src/batched_solver.cu:122-178)
template <bool CHECK, int NG>
struct BXOp : BOpBase<NG> {
    using Base = BOpBase<NG>;
    const double *Y;
    double *X, *X_hat;
    const double *L, *U, *C, *lastX;
    double *DX, *Z_bar, *X_bar;
    const double *sigma;
    const int *kx;
    int *ky;
    const unsigned char *active;
    double sig[NG], f1[NG], f2[NG];
    bool on[NG];
    __device__ __forceinline__ void init(int g0, int lane, int G) {
        Base::init_base(Y, g0, lane, G);
#pragma unroll
        for (int q = 0; q < NG; ++q) {
            const int inst = (Base::valid[q] ? g0 + q : g0) * kGS + lane;
            sig[q] = sigma[inst];
            on[q] = Base::valid[q] && active[inst] != 0;
            const int k = kx[inst];
            f1[q] = 1.0 / (k + 2.0);
            f2[q] = 1.0 - f1[q];
            if (Base::valid[q] && blockIdx.x == 0 && threadIdx.x < 32) ky[inst] = k;
        }
    }
    struct Pre { double xi, c, l, u, x0; };
    __device__ __forceinline__ Pre pre(int r, int q) const {
        Pre p{0.0, 0.0, 0.0, 0.0, 0.0};
        if (!on[q]) return p;
        const size_t t = Base::obase[q] + (size_t)r * kGS;
        p.xi = ld_once(X + t); p.c = ld_once(C + t); p.l = ld_once(L + t); p.u = ld_once(U + t); p.x0 = ld_once(lastX + t);
        return p;
    }
    __device__ __forceinline__ void row(int r, int q, double acc, const Pre &p) const {
        if (!on[q]) return;
        const size_t t = Base::obase[q] + (size_t)r * kGS;
        const double xi = p.xi;
        const double zt = fma(sig[q], acc - p.c, xi);
        const double xb = fmin(fmax(zt, p.l), p.u);
        const double xh = 2.0 * xb - xi;
        if (CHECK) {
            st_once(DX + t, xb - xh);
            st_once(Z_bar + t, (xb - zt) / sig[q]);
            st_once(X_bar + t, xb);
        }
        X_hat[t] = xh;   // gathered by the y-phase that follows: normal priority
        st_once(X + t, fma(f2[q], xh, f1[q] * p.x0));
    }
};

// y-phase (reference update_y_{normal,check}_batched_kernel :180-236): y_bar = d / (lambda sigma) (a division
// here, a reciprocal multiply in the single-instance path -- reference quirk #5)
template <bool CHECK, int NG>
struct BYOp : BOpBase<NG> {
    using Base = BOpBase<NG>;
    const double *X_hat;
    double *Y;
    const double *AL, *AU, *lastY;
    double *DY, *Y_bar, *Y_obj;
    const double *sigma;
    const int *ky;
    int *kx;
    const unsigned char *active;
    double lambda_max;
    double fact1[NG], f1[NG], f2[NG];
    bool on[NG];
    __device__ __forceinline__ void init(int g0, int lane, int G) {
        Base::init_base(X_hat, g0, lane, G);
#pragma unroll
        for (int q = 0; q < NG; ++q) {
            const int inst = (Base::valid[q] ? g0 + q : g0) * kGS + lane;
            fact1[q] = lambda_max * sigma[inst];
            on[q] = Base::valid[q] && active[inst] != 0;
            const int k = ky[inst];
            f1[q] = 1.0 / (k + 2.0);
            f2[q] = 1.0 - f1[q];
            if (on[q] && blockIdx.x == 0 && threadIdx.x < 32) kx[inst] = k + 1;
        }
    }
    struct Pre { double yi, al, au, y0; };
    __device__ __forceinline__ Pre pre(int r, int q) const {
        Pre p{0.0, 0.0, 0.0, 0.0};
        if (!on[q]) return p;
        const size_t t = Base::obase[q] + (size_t)r * kGS;
        p.yi = ld_once(Y + t); p.al = ld_once(AL + t); p.au = ld_once(AU + t); p.y0 = ld_once(lastY + t);
        return p;
    }
    __device__ __forceinline__ void row(int r, int q, double acc, const Pre &p) const {
        if (!on[q]) return;
        const size_t t = Base::obase[q] + (size_t)r * kGS;
        const double yi = p.yi;
        const double v = fma(-fact1[q], yi, acc);
        const double d = fmax(p.al - v, fmin(p.au - v, 0.0));
        const double yb = d / fact1[q];
        const double yh = 2.0 * yb - yi;
        if (CHECK) {
            st_once(DY + t, yb - yh);
            st_once(Y_bar + t, yb);
            st_once(Y_obj + t, v + d);
        }
        Y[t] = fma(f2[q], yh, f1[q] * p.y0);   // gathered by the next x-phase: normal priority
    }
};

// dual residual + objective terms (reference compute_batched_Rd_kernel :238-249 + cublasDdot/Dnrm2 :604-607,
// lu violation :265-278,615-617): slots 0 |RD|^2, 1 <C,X_bar>, 2 <X_bar,Z_bar>, 3 |lu/col_norm|^2 (iter 0)
template <bool ITER0>
struct BResDualOp : BOpBase<1> {
    const double *Y_bar, *C, *Z_bar, *X_bar, *L, *U, *col_norm;
    double *partials;
    double t[4];
    __device__ __forceinline__ void init(int g0, int lane, int G) { init_base(Y_bar, g0, lane, G); t[0] = t[1] = t[2] = t[3] = 0.0; }
    __device__ __forceinline__ void row(int j, int, double acc, const Pre &) {
        const size_t i = obase[0] + (size_t)j * kGS;
        const double cj = C[i], zb = Z_bar[i], xb = X_bar[i], cn = col_norm[j];
        const double rd = (cj - acc - zb) * cn;
        t[0] += rd * rd;
        t[1] += cj * xb;
        t[2] += xb * zb;
        if (ITER0) {
            const double lo = L[i], hi = U[i];
            const double viol = xb < lo ? lo - xb : (xb > hi ? xb - hi : 0.0);
            const double q = viol / cn;
            t[3] += q * q;
        }
    }
    __device__ __forceinline__ void finish(double (*red)[kMaxSlots][kGS], int warp, int lane) {
        batched_reduce_store<4>(t, red, warp, lane, partials);
    }
};

// primal residual (reference compute_batched_Rp_kernel :251-263 + :605,608): slots 0 |RP|^2, 1 <Y_obj,Y_bar>
struct BResPrimalOp : BOpBase<1> {
    const double *X_bar, *AL, *AU, *row_norm, *Y_obj, *Y_bar;
    double *partials;
    double t[2];
    __device__ __forceinline__ void init(int g0, int lane, int G) { init_base(X_bar, g0, lane, G); t[0] = t[1] = 0.0; }
    __device__ __forceinline__ void row(int r, int, double ax, const Pre &) {
        const size_t i = obase[0] + (size_t)r * kGS;
        const double rp = row_norm[r] * fmax(fmin(AU[i] - ax, 0.0), AL[i] - ax);
        t[0] += rp * rp;
        t[1] += Y_obj[i] * Y_bar[i];
    }
    __device__ __forceinline__ void finish(double (*red)[kMaxSlots][kGS], int warp, int lane) {
        batched_reduce_store<2>(t, red, warp, lane, partials);
    }
};

// M-norm terms (reference compute_weighted_norm :625-650): slots 0 <A DX, DY>, 1 |DY|^2
struct BWeightedOp : BOpBase<1> {
    const double *DX, *DY;
    double *partials;
    double t[2];
    __device__ __forceinline__ void init(int g0, int lane, int G) { init_base(DX, g0, lane, G); t[0] = t[1] = 0.0; }
    __device__ __forceinline__ void row(int r, int, double acc, const Pre &) {
        const double dy = DY[obase[0] + (size_t)r * kGS];
        t[0] += acc * dy;
        t[1] += dy * dy;
    }
    __device__ __forceinline__ void finish(double (*red)[kMaxSlots][kGS], int warp, int lane) {
        batched_reduce_store<2>(t, red, warp, lane, partials);
    }
};

// ---------------------------------------------------------------------------------------------
// dense [group][row][32] vector kernels: grid (chunks, groups), warp per row
// ---------------------------------------------------------------------------------------------
// per-instance |V|^2
__global__ void __launch_bounds__(kBThreads) batched_sumsq_kernel(const double *V, int rows, double *partials) {
    __shared__ double red[kBWarps][kMaxSlots][kGS];
    const int g = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double t[1] = {0.0};
    for (int r = blockIdx.x * kBWarps + warp; r < rows; r += gridDim.x * kBWarps) {
        const double v = V[((size_t)g * rows + r) * kGS + lane];
        t[0] += v * v;
    }
    batched_reduce_store<1>(t, red, warp, lane, partials);
}

// movement norms (reference batched_restart_movement_kernel :280-294 + nrm2 :661-664) fused with the masked
// restart copy (do_batched_restart_kernel :296-323): slots 0 |X_bar-lastX|^2, 1 |Y_bar-lastY|^2
__global__ void __launch_bounds__(kBThreads) batched_restart_kernel(const double *X_bar, double *lastX, double *X, int n, const double *Y_bar,
                                                                   double *lastY, double *Y, int m, const unsigned char *flags,
                                                                   int *kx, double *partials) {
    __shared__ double red[kBWarps][kMaxSlots][kGS];
    const int g = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool flag = flags[g * kGS + lane] != 0;
    double t[2] = {0.0, 0.0};
    for (int r = blockIdx.x * kBWarps + warp; r < n; r += gridDim.x * kBWarps) {
        const size_t i = ((size_t)g * n + r) * kGS + lane;
        const double xb = X_bar[i], d = xb - lastX[i];
        t[0] += d * d;
        if (flag) { lastX[i] = xb; X[i] = xb; }
    }
    for (int r = blockIdx.x * kBWarps + warp; r < m; r += gridDim.x * kBWarps) {
        const size_t i = ((size_t)g * m + r) * kGS + lane;
        const double yb = Y_bar[i], d = yb - lastY[i];
        t[1] += d * d;
        if (flag) { lastY[i] = yb; Y[i] = yb; }
    }
    if (blockIdx.x == 0 && warp == 0 && flag) kx[g * kGS + lane] = 0;   // Halpern counter reset for restarted instances
    batched_reduce_store<2>(t, red, warp, lane, partials);
}

// column-major host layout (instance k, row i at k*rows + i) <-> device layout [g][i][32]
// device layout -> column-major with unscaling: out = (v (/|*) norm[row]) * scale[instance]
// (reference collect_results :915-924)
template <bool DIVIDE>
__global__ void from_group_layout_kernel(const double *src, double *dst, int rows, int B, const double *norm, const double *scale) {
    __shared__ double tile[32][33];
    const int g = blockIdx.y, r0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int k = ty; k < 32; k += 8) {
        const int r = r0 + k;
        if (r < rows) {
            const double v = src[((size_t)g * rows + r) * kGS + tx];
            const double nr = norm[r];
            tile[k][tx] = (DIVIDE ? v / nr : v * nr) * scale[g * kGS + tx];
        }
    }
    __syncthreads();
    for (int k = ty; k < 32; k += 8) {
        const int inst = g * kGS + k, r = r0 + tx;
        if (inst < B && r < rows) dst[(size_t)inst * rows + r] = tile[tx][k];
    }
}

// ---------------------------------------------------------------------------------------------
// per-instance scaling on the DEVICE (reference build_batched_lp_device, src/batched_solver.cu:792-885, runs these
// loops on the host, one instance after the other).  The inputs arrive column-major (instance k, row i at k*rows + i),
// exactly as the caller holds them; every elementwise operation is the reference's (same divisions, same order, two
// separately rounded steps).  The reference accumulates its norms in long double (:332-354); here they are accumulated
// in double-double (error-free products and sums, ~106 bits) in a fixed order, i.e. at least as accurately, and rounded
// to double once -- the result can differ from the reference's by one ulp of the norm when its own 64-bit accumulation
// error crosses a rounding boundary.
// ---------------------------------------------------------------------------------------------
struct DD { double hi, lo; };
__device__ __forceinline__ DD dd_add(DD a, DD b) {
    const double s = a.hi + b.hi;
    const double bb = s - a.hi;
    const double e = (a.hi - (s - bb)) + (b.hi - bb);
    const double t = e + (a.lo + b.lo);
    const double hi = s + t;
    return DD{hi, t - (hi - s)};
}
__device__ __forceinline__ DD dd_add_sq(DD a, double v) {
    const double p = v * v;
    return dd_add(a, DD{p, fma(v, v, -p)});
}
__device__ __forceinline__ double bound_abs(double lo, double hi) {   // reference bound_norm_host :332-343
    const double a = (isinf(lo) && lo < 0) ? 0.0 : fabs(lo);
    const double b = (isinf(hi) && hi > 0) ? 0.0 : fabs(hi);
    return fmax(a, b);
}
constexpr int kNormChunks = 32;   // CTAs per instance in the norm passes
// fixed-order CTA reduction of NS double-double accumulators; thread 0 writes partial[(inst*kNormChunks + chunk)*NS + s]
template <int NS>
__device__ __forceinline__ void dd_block_store(DD (&acc)[NS], DD *partial) {
    __shared__ DD sm[NS][kBWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        DD v = acc[s];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            DD o;
            o.hi = __shfl_xor_sync(0xffffffffu, v.hi, off);
            o.lo = __shfl_xor_sync(0xffffffffu, v.lo, off);
            v = dd_add(v, o);
        }
        if (lane == 0) sm[s][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            DD v = sm[s][0];
            for (int w = 1; w < kBWarps; ++w) v = dd_add(v, sm[s][w]);
            partial[((size_t)blockIdx.y * kNormChunks + blockIdx.x) * NS + s] = v;
        }
    }
}
// rows: slot 0 |b|^2 of the raw bounds, then AL,AU /= row_norm in place, slot 1 |b|^2 of the scaled bounds
__global__ void __launch_bounds__(kBThreads) batched_scale_rows_kernel(double *AL, double *AU, const double *row_norm, int m, DD *partial) {
    const size_t base = (size_t)blockIdx.y * m;
    DD acc[2] = {{0.0, 0.0}, {0.0, 0.0}};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += kNormChunks * blockDim.x) {
        double lo = AL[base + i], hi = AU[base + i];
        acc[0] = dd_add_sq(acc[0], bound_abs(lo, hi));
        const double rn = row_norm[i];
        lo = lo / rn; hi = hi / rn;
        AL[base + i] = lo; AU[base + i] = hi;
        acc[1] = dd_add_sq(acc[1], bound_abs(lo, hi));
    }
    dd_block_store<2>(acc, partial);
}
// columns: slot 0 |c|^2 raw, then C /= col_norm in place, slot 1 |c|^2 scaled
__global__ void __launch_bounds__(kBThreads) batched_scale_cols_kernel(double *Cm, const double *col_norm, int n, DD *partial) {
    const size_t base = (size_t)blockIdx.y * n;
    DD acc[2] = {{0.0, 0.0}, {0.0, 0.0}};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += kNormChunks * blockDim.x) {
        double c = Cm[base + i];
        acc[0] = dd_add_sq(acc[0], c);
        c = c / col_norm[i];
        Cm[base + i] = c;
        acc[1] = dd_add_sq(acc[1], c);
    }
    dd_block_store<2>(acc, partial);
}
// second pass (bounds/cost scaling): divide by the instance's scale in place (if on), slot 0 = norm^2 of the result
__global__ void __launch_bounds__(kBThreads) batched_div_rows_kernel(double *AL, double *AU, const double *scale, bool on, int m, DD *partial) {
    const size_t base = (size_t)blockIdx.y * m;
    const double sc = scale[blockIdx.y];
    DD acc[1] = {{0.0, 0.0}};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += kNormChunks * blockDim.x) {
        double lo = AL[base + i], hi = AU[base + i];
        if (on) { lo = lo / sc; hi = hi / sc; AL[base + i] = lo; AU[base + i] = hi; }
        acc[0] = dd_add_sq(acc[0], bound_abs(lo, hi));
    }
    dd_block_store<1>(acc, partial);
}
__global__ void __launch_bounds__(kBThreads) batched_div_cols_kernel(double *Cm, const double *scale, bool on, int n, DD *partial) {
    const size_t base = (size_t)blockIdx.y * n;
    const double sc = scale[blockIdx.y];
    DD acc[1] = {{0.0, 0.0}};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += kNormChunks * blockDim.x) {
        double c = Cm[base + i];
        if (on) { c = c / sc; Cm[base + i] = c; }
        acc[0] = dd_add_sq(acc[0], c);
    }
    dd_block_store<1>(acc, partial);
}
// out[s][k] = sqrt(sum over the instance's chunks, in chunk order) (+ 1 where plus_one[s]); on == false: out = 1
__global__ void batched_norm_finalize_kernel(const DD *partial, int ns, int B, int Bpad, double *out, int plus_one_mask, int skip_mask) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= B) return;
    for (int s = 0; s < ns; ++s) {
        DD v = partial[((size_t)k * kNormChunks) * ns + s];
        for (int c = 1; c < kNormChunks; ++c) v = dd_add(v, partial[((size_t)k * kNormChunks + c) * ns + s]);
        double r = sqrt(v.hi + v.lo);
        if (plus_one_mask & (1 << s)) r = 1.0 + r;
        if (skip_mask & (1 << s)) r = 1.0;
        out[(size_t)s * Bpad + k] = r;
    }
}
// column-major -> [g][row][32] with the last elementwise steps fused:
//   COLSCALE (l, u): v *= col_norm[row], then (bc) v /= b_scale[instance]          (reference :835-847)
//   REPL -1: -inf -> -1e100 (AL, l);  +1: +inf -> +1e100 (AU, u);  0: none (C)      (reference :849-864)
template <int REPL, bool COLSCALE>
__global__ void to_group_layout_scaled_kernel(const double *src, double *dst, int rows, int B, const double *col_norm,
                                              const double *b_scale, bool bc) {
    __shared__ double tile[32][33];
    const int g = blockIdx.y, r0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    for (int k = ty; k < 32; k += 8) {
        const int inst = g * kGS + k, r = r0 + tx;
        double v = 0.0;
        if (inst < B && r < rows) {
            v = src[(size_t)inst * rows + r];
            if (COLSCALE) {
                v = v * col_norm[r];
                if (bc) v = v / b_scale[inst];
            }
            if (REPL < 0 && isinf(v) && v < 0) v = -kInfReplacement;
            if (REPL > 0 && isinf(v) && v > 0) v = kInfReplacement;
        }
        tile[k][tx] = v;
    }
    __syncthreads();
    for (int k = ty; k < 32; k += 8) {
        const int r = r0 + k;
        if (r < rows) dst[((size_t)g * rows + r) * kGS + tx] = tile[tx][k];
    }
}

// row-major (rows x B, numpy's default order for an (n, B) array) -> column-major staging (instance k, row i at k*rows + i):
// what the reference's Python binding does element by element on the host (bindings/python/src/hprlp_pybind.cpp:343-356)
__global__ void rowmajor_to_colmajor_kernel(const double *src, double *dst, int rows, int B) {
    __shared__ double tile[32][33];
    const int r0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    for (int j = ty; j < 32; j += 8) {
        const int r = r0 + j, k = k0 + tx;
        if (r < rows && k < B) tile[j][tx] = src[(size_t)r * B + k];
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        const int k = k0 + j, r = r0 + tx;
        if (r < rows && k < B) dst[(size_t)k * rows + r] = tile[tx][j];
    }
}

// pinned host blocks recycled for the life of the process (cudaFreeHost synchronises the device: engine.cu)
std::mutex g_bpin_mu;
std::vector<std::pair<double *, size_t>> g_bpin_free;
double *bpinned_acquire(size_t doubles) {
    {
        std::lock_guard<std::mutex> lk(g_bpin_mu);
        for (size_t i = 0; i < g_bpin_free.size(); ++i)
            if (g_bpin_free[i].second >= doubles) { double *p = g_bpin_free[i].first; g_bpin_free.erase(g_bpin_free.begin() + i); return p; }
    }
    double *p = nullptr;
    HPR_CUDA_CHECK(cudaHostAlloc(reinterpret_cast<void **>(&p), sizeof(double) * doubles, cudaHostAllocPortable));
    return p;
}
void bpinned_release(double *p, size_t doubles) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_bpin_mu);
    g_bpin_free.emplace_back(p, doubles);
}

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int step_of(int iter) { return std::max(10, static_cast<int>(pow(10, floor(log10((double)iter))) / 10)); }

HPRLP_batched_results make_batched_error(const char *status, int m, int n, int B) {   // reference :356-368
    HPRLP_batched_results r;
    r.m = m; r.n = n; r.batch_size = B;
    if (B > 0) {
        r.status = static_cast<char *>(std::calloc(static_cast<size_t>(B) * 64, sizeof(char)));
        for (int k = 0; k < B; ++k) std::strncpy(r.status + 64 * k, status, 63);
    }
    return r;
}

struct RestartHost {   // reference BatchedRestartHost :103-120
    std::vector<int> restart_flag, inner, times;
    std::vector<unsigned char> first_restart;
    std::vector<double> last_gap, current_gap, save_gap, best_gap, best_sigma, sigma;
};

struct ResidualHost {   // reference BatchedResidualHost :94-101
    std::vector<double> primal_obj, dual_obj, err_Rp, err_Rd, rel_gap, kkt_error;
};

class BatchedSolver {
   public:
    Engine eng;   // shared matrix: upload, scaling (bc off), power iteration
    int m = 0, n = 0, B = 0, G = 0, Bpad = 0;
    double *X = nullptr, *X_hat = nullptr, *X_bar = nullptr, *DX = nullptr, *Z_bar = nullptr, *lastX = nullptr;
    double *C = nullptr, *L = nullptr, *U = nullptr;
    double *Y = nullptr, *Y_bar = nullptr, *DY = nullptr, *Y_obj = nullptr, *lastY = nullptr, *AL = nullptr, *AU = nullptr;
    double *d_sigma = nullptr, *d_scale_b = nullptr, *d_scale_c = nullptr;
    int *d_k = nullptr;   // [kx (Bpad), ky (Bpad)]
    unsigned char *d_active = nullptr, *d_flags = nullptr;
    double *d_partials = nullptr, *d_scal = nullptr, *h_scal = nullptr;
    double *d_stage = nullptr;   // column-major staging buffer, max(n,m) * B
    double lambda_max = 1.0;
    cudaStream_t stream = nullptr;
    int nbx_A = 0, nbx_AT = 0, nbx_vec = 0;
    long long launches = 0;
    std::vector<double> b_scale, c_scale, norm_b, norm_c, norm_b_org, norm_c_org, obj_constants;

    void *arena = nullptr;       // one block of the engines' pool behind every device array below
    size_t h_scal_doubles = 0;
    DD *d_dd = nullptr;          // double-double partials of the setup norms
    double *d_stage2 = nullptr;  // second / third column-major staging buffers (m * B each) for AL, AU
    double *d_stage3 = nullptr;
    double *d_norms = nullptr;   // [6][Bpad]: norm_b_org, b_scale, norm_c_org, c_scale, norm_b, norm_c
    double *d_raw = nullptr;     // row-major inputs as uploaded (layout 1 only)

    ~BatchedSolver() {
        if (arena) pool_free(arena, stream);
        bpinned_release(h_scal, h_scal_doubles);
    }

    BView viewA() const { return BView{m, n, eng.A.rowPtr, eng.A.col, eng.A.val}; }
    BView viewAT() const { return BView{n, m, eng.AT.rowPtr, eng.AT.col, eng.AT.val}; }
    dim3 gridA(int ng = 1) const { return dim3(nbx_A, (G + ng - 1) / ng); }
    dim3 gridAT(int ng = 1) const { return dim3(nbx_AT, (G + ng - 1) / ng); }
    // Groups of 32 instances per pass of the x- / y-phase (HPRLP_BATCH_NGX / _NGY: 1, 2 or 4).  Measured on configs[3]
    // (B200, 300 iterations, profiles/r2_batched_group_sweep.md): (1,1) 421 ms, (2,1) 429, (4,1) 481, (2,2) 453, (4,2) 504,
    // (4,4) 591 -- more groups per pass means more gathered slabs competing for the L2, and the passes are bound by L2
    // throughput on the 256-byte gathers (ncu: ~9-10 TB/s through the L2 against ~4 TB/s from DRAM), not by issue.
    int ngx = 1, ngy = 1;

    void fetch(int slots) {
        HPR_CUDA_CHECK(cudaMemcpyAsync(h_scal, d_scal, sizeof(double) * (size_t)slots * Bpad, cudaMemcpyDeviceToHost, stream));
        HPR_CUDA_CHECK(cudaStreamSynchronize(stream));
    }
    double hs(int slot, int k) const { return h_scal[(size_t)slot * Bpad + k]; }

    template <bool CHECK, int NG>
    void launch_x() {
        BXOp<CHECK, NG> ox{};
        ox.gcols = m; ox.orows = n;
        ox.Y = Y; ox.X = X; ox.X_hat = X_hat; ox.L = L; ox.U = U; ox.C = C; ox.lastX = lastX;
        ox.DX = DX; ox.Z_bar = Z_bar; ox.X_bar = X_bar; ox.sigma = d_sigma; ox.kx = d_k; ox.ky = d_k + Bpad; ox.active = d_active;
        batched_rows_kernel<<<gridAT(NG), kBThreads, 0, stream>>>(viewAT(), ox, G);
    }
    template <bool CHECK, int NG>
    void launch_y() {
        BYOp<CHECK, NG> oy{};
        oy.gcols = n; oy.orows = m;
        oy.X_hat = X_hat; oy.Y = Y; oy.AL = AL; oy.AU = AU; oy.lastY = lastY; oy.DY = DY; oy.Y_bar = Y_bar; oy.Y_obj = Y_obj;
        oy.sigma = d_sigma; oy.ky = d_k + Bpad; oy.kx = d_k; oy.active = d_active; oy.lambda_max = lambda_max;
        batched_rows_kernel<<<gridA(NG), kBThreads, 0, stream>>>(viewA(), oy, G);
    }
    void iteration(bool check) {
        if (check) { launch_x<true, 1>(); launch_y<true, 1>(); }   // 1 in 10 iterations at most: one instantiation each
        else {
            if (ngx >= 4) launch_x<false, 4>(); else if (ngx == 2) launch_x<false, 2>(); else launch_x<false, 1>();
            if (ngy >= 4) launch_y<false, 4>(); else if (ngy == 2) launch_y<false, 2>(); else launch_y<false, 1>();
        }
        launches += 2;
    }

    // reference compute_weighted_norm :625-650 (lambda_max shared by the batch, only ever increased)
    std::vector<double> weighted_norm(const std::vector<double> &sigma) {
        BWeightedOp o{}; o.gcols = n; o.orows = m; o.DX = DX; o.DY = DY; o.partials = d_partials;
        batched_rows_kernel<<<gridA(), kBThreads, 0, stream>>>(viewA(), o, G);
        batched_final_reduce_kernel<<<G, 32 * 2, 0, stream>>>(d_partials, nbx_A, 2, Bpad, d_scal);
        batched_sumsq_kernel<<<dim3(nbx_vec, G), kBThreads, 0, stream>>>(DX, n, d_partials);
        batched_final_reduce_kernel<<<G, 32, 0, stream>>>(d_partials, nbx_vec, 1, Bpad, d_scal + 2 * (size_t)Bpad);
        launches += 4;
        fetch(3);
        std::vector<double> w(B, 0.0);
        for (int k = 0; k < B; ++k) {
            const double dot_prod = 2.0 * hs(0, k), dy_sq = hs(1, k), dx_sq = hs(2, k);
            double value = sigma[k] * (lambda_max * dy_sq) + dx_sq / sigma[k] + dot_prod;
            if (value < 0.0 && dy_sq > 0.0) {
                const double cand = -(dot_prod + dx_sq / sigma[k]) / (sigma[k] * dy_sq) * 1.05;
                lambda_max = std::max(lambda_max, cand);
                value = sigma[k] * (lambda_max * dy_sq) + dx_sq / sigma[k] + dot_prod;
            }
            w[k] = std::sqrt(std::max(value, 0.0));
        }
        return w;
    }

    // reference compute_residuals :578-623
    void residuals(int iter, ResidualHost *res) {
        if (iter == 0) {
            BResDualOp<true> o{}; o.gcols = m; o.orows = n; o.Y_bar = Y_bar; o.C = C; o.Z_bar = Z_bar; o.X_bar = X_bar; o.L = L; o.U = U;
            o.col_norm = eng.col_norm; o.partials = d_partials;
            batched_rows_kernel<<<gridAT(), kBThreads, 0, stream>>>(viewAT(), o, G);
        } else {
            BResDualOp<false> o{}; o.gcols = m; o.orows = n; o.Y_bar = Y_bar; o.C = C; o.Z_bar = Z_bar; o.X_bar = X_bar; o.L = L; o.U = U;
            o.col_norm = eng.col_norm; o.partials = d_partials;
            batched_rows_kernel<<<gridAT(), kBThreads, 0, stream>>>(viewAT(), o, G);
        }
        batched_final_reduce_kernel<<<G, 32 * 4, 0, stream>>>(d_partials, nbx_AT, 4, Bpad, d_scal);
        BResPrimalOp p{}; p.gcols = n; p.orows = m; p.X_bar = X_bar; p.AL = AL; p.AU = AU; p.row_norm = eng.row_norm; p.Y_obj = Y_obj; p.Y_bar = Y_bar;
        p.partials = d_partials;
        batched_rows_kernel<<<gridA(), kBThreads, 0, stream>>>(viewA(), p, G);
        batched_final_reduce_kernel<<<G, 32 * 2, 0, stream>>>(d_partials, nbx_A, 2, Bpad, d_scal + 4 * (size_t)Bpad);
        launches += 4;
        fetch(6);
        for (int k = 0; k < B; ++k) {
            const double obj_scale = b_scale[k] * c_scale[k];
            res->primal_obj[k] = obj_scale * hs(1, k) + obj_constants[k];
            res->dual_obj[k] = obj_scale * (hs(5, k) + hs(2, k)) + obj_constants[k];
            res->err_Rd[k] = c_scale[k] * std::sqrt(hs(0, k)) / norm_c_org[k];
            res->err_Rp[k] = b_scale[k] * std::sqrt(hs(4, k)) / norm_b_org[k];
            if (iter == 0) res->err_Rp[k] = std::max(res->err_Rp[k], b_scale[k] * std::sqrt(hs(3, k)));
            res->rel_gap[k] = std::abs(res->primal_obj[k] - res->dual_obj[k]) /
                              (1.0 + std::abs(res->primal_obj[k]) + std::abs(res->dual_obj[k]));
            res->kkt_error[k] = std::max(res->err_Rp[k], std::max(res->err_Rd[k], res->rel_gap[k]));
        }
    }
};

// reference check_restart :667-700
void check_restart(RestartHost *r, int iter, int check_iter, const std::vector<unsigned char> &active) {
    for (int k = 0; k < (int)active.size(); ++k) {
        if (!active[k]) continue;
        if (r->first_restart[k]) {
            if (iter == check_iter) {
                r->first_restart[k] = 0;
                r->restart_flag[k] = 1;
                r->best_gap[k] = r->current_gap[k];
                r->best_sigma[k] = r->sigma[k];
            }
        } else if (iter % check_iter == 0) {
            if (r->current_gap[k] < 0.0) r->current_gap[k] = 1.0e-6;
            if (r->current_gap[k] <= 0.2 * r->last_gap[k]) r->restart_flag[k] = 1;
            if (r->current_gap[k] <= 0.6 * r->last_gap[k] && r->current_gap[k] > r->save_gap[k]) r->restart_flag[k] = 2;
            if (r->inner[k] >= 0.2 * iter) r->restart_flag[k] = 3;
            if (r->best_gap[k] > r->current_gap[k]) { r->best_gap[k] = r->current_gap[k]; r->best_sigma[k] = r->sigma[k]; }
            r->save_gap[k] = r->current_gap[k];
        }
    }
}

}  // namespace
}  // namespace hpr

using namespace hpr;

extern "C" void free_batched_results(HPRLP_batched_results *results);

// One GPU: the whole batch on device param->device_number.
// layout 0: inputs column-major (the ABI: instance k contiguous); 1: row-major rows x B (C-ordered numpy (n, B) arrays)
static HPRLP_batched_results solve_batched_on_device(const LP_info_cpu *model, int batch_size, const HPRLP_FLOAT *C_in,
                                                     const HPRLP_FLOAT *AL_in, const HPRLP_FLOAT *AU_in, const HPRLP_FLOAT *l_in,
                                                     const HPRLP_FLOAT *u_in, const HPRLP_FLOAT *obj_constants,
                                                     const HPRLP_parameters *param, int layout = 0) {
    HPRLP_parameters def;
    HPRLP_parameters actual = param ? *param : def;
    actual.use_presolve = false;
    const double setup_start = now_s();
    const int m = model->m, n = model->n, B = batch_size;

    BatchedSolver S;
    S.m = m; S.n = n; S.B = B; S.G = (B + kGS - 1) / kGS; S.Bpad = S.G * kGS;
    // shared matrix on the device with dummy vectors; matrix-only scaling (bc off) -- reference :959-989
    {
        std::vector<double> zero_m(m, 0.0), zero_n(n, 0.0);
        LP_info_cpu mat{};
        mat.m = m; mat.n = n; mat.A = model->A;
        mat.AL = zero_m.data(); mat.AU = zero_m.data(); mat.c = zero_n.data(); mat.l = zero_n.data(); mat.u = zero_n.data();
        mat.obj_constant = 0.0;
        S.eng.upload(&mat, actual.device_number);
        HPRLP_parameters mp = actual;
        mp.use_bc_scaling = false;
        S.eng.scale(&mp);
    }
    S.stream = S.eng.stream;
    if (const char *e = getenv("HPRLP_BATCH_NGX")) S.ngx = atoi(e);
    if (const char *e = getenv("HPRLP_BATCH_NGY")) S.ngy = atoi(e);
    cudaStream_t st = S.stream;
    const bool bc = actual.use_bc_scaling;
    static const bool timing = getenv("HPRLP_TIMING") != nullptr;   // stage wall times on stderr
    double t_mark = setup_start;
    auto stage_done = [&](const char *what) {
        if (!timing) return;
        cudaStreamSynchronize(st);
        const double t = now_s();
        std::fprintf(stderr, "[hprlp timing] batched %s %.4f s\n", what, t - t_mark);
        t_mark = t;
    };
    stage_done("matrix upload + scaling");

    // device state: one zero-filled block of the engines' pool, carved into the arrays
    const size_t nG = (size_t)n * S.Bpad, mG = (size_t)m * S.Bpad;
    S.nbx_A = (m + kRowsPerCta - 1) / kRowsPerCta;
    S.nbx_AT = (n + kRowsPerCta - 1) / kRowsPerCta;
    S.nbx_vec = std::max(1, std::min((std::max(m, n) + kBWarps - 1) / kBWarps, 148 * 2));
    const size_t nblk = (size_t)std::max(std::max(S.nbx_A, S.nbx_AT), S.nbx_vec) * S.G;
    {
        auto up = [](size_t b) { return (b + 511) & ~(size_t)511; };
        size_t off = 0;
        auto take = [&](size_t bytes) { const size_t o = off; off += up(bytes); return o; };
        const size_t o_n[9] = {take(nG * 8), take(nG * 8), take(nG * 8), take(nG * 8), take(nG * 8), take(nG * 8), take(nG * 8), take(nG * 8), take(nG * 8)};
        const size_t o_m[7] = {take(mG * 8), take(mG * 8), take(mG * 8), take(mG * 8), take(mG * 8), take(mG * 8), take(mG * 8)};
        const size_t o_sigma = take(S.Bpad * 8), o_sb = take(S.Bpad * 8), o_sc = take(S.Bpad * 8), o_k = take(2 * (size_t)S.Bpad * 4);
        const size_t o_act = take(S.Bpad), o_flg = take(S.Bpad);
        const size_t o_part = take(nblk * kMaxSlots * kGS * 8), o_scal = take((size_t)kMaxSlots * S.Bpad * 8);
        const size_t o_stage = take((size_t)n * B * 8), o_stage2 = take((size_t)m * B * 8), o_stage3 = take((size_t)m * B * 8);
        const size_t o_dd = take((size_t)B * kNormChunks * 2 * sizeof(DD)), o_norms = take((size_t)6 * S.Bpad * 8);
        const size_t o_raw = take(layout == 1 ? (size_t)n * B * 8 : 8);
        char *base = static_cast<char *>(pool_alloc_zeroed(off, actual.device_number, st));
        S.arena = base;
        double **nv[9] = {&S.X, &S.X_hat, &S.X_bar, &S.DX, &S.Z_bar, &S.lastX, &S.C, &S.L, &S.U};
        for (int i = 0; i < 9; ++i) *nv[i] = reinterpret_cast<double *>(base + o_n[i]);
        double **mv[7] = {&S.Y, &S.Y_bar, &S.DY, &S.Y_obj, &S.lastY, &S.AL, &S.AU};
        for (int i = 0; i < 7; ++i) *mv[i] = reinterpret_cast<double *>(base + o_m[i]);
        S.d_sigma = reinterpret_cast<double *>(base + o_sigma); S.d_scale_b = reinterpret_cast<double *>(base + o_sb);
        S.d_scale_c = reinterpret_cast<double *>(base + o_sc); S.d_k = reinterpret_cast<int *>(base + o_k);
        S.d_active = reinterpret_cast<unsigned char *>(base + o_act); S.d_flags = reinterpret_cast<unsigned char *>(base + o_flg);
        S.d_partials = reinterpret_cast<double *>(base + o_part); S.d_scal = reinterpret_cast<double *>(base + o_scal);
        S.d_stage = reinterpret_cast<double *>(base + o_stage); S.d_stage2 = reinterpret_cast<double *>(base + o_stage2);
        S.d_stage3 = reinterpret_cast<double *>(base + o_stage3);
        S.d_dd = reinterpret_cast<DD *>(base + o_dd); S.d_norms = reinterpret_cast<double *>(base + o_norms);
        S.d_raw = reinterpret_cast<double *>(base + o_raw);
    }
    S.h_scal_doubles = (size_t)kMaxSlots * S.Bpad;
    S.h_scal = bpinned_acquire(S.h_scal_doubles);
    stage_done("state allocation");

    // The power iteration needs the shared, scaled matrix only (reference :994-1001 runs it after the instances are built):
    // it runs on the engine's stream from a helper thread while this thread stages and scales the instances on a second
    // stream.  On configs[3] both take 0.04-0.06 s and the iteration's small kernels leave the copy engines and most SMs idle.
    struct UploadStream {
        cudaStream_t s = nullptr;
        cudaEvent_t ready = nullptr;
        ~UploadStream() { if (ready) cudaEventDestroy(ready); if (s) cudaStreamDestroy(s); }
    } upl;
    HPR_CUDA_CHECK(cudaStreamCreateWithFlags(&upl.s, cudaStreamNonBlocking));
    HPR_CUDA_CHECK(cudaEventCreateWithFlags(&upl.ready, cudaEventDisableTiming));
    HPR_CUDA_CHECK(cudaEventRecord(upl.ready, st));          // matrix scaling and the arena's zero-fill are queued on st
    HPR_CUDA_CHECK(cudaStreamWaitEvent(upl.s, upl.ready, 0));
    double power_time = 0.0;
    std::exception_ptr power_error;
    std::thread power_thread([&]() {
        try {
            HPR_CUDA_CHECK(cudaSetDevice(actual.device_number));
            const double t0 = now_s();
            S.lambda_max = S.eng.power_iteration(5000, 1.0e-4, nullptr, nullptr) * 1.01;
            power_time = now_s() - t0;
        } catch (...) { power_error = std::current_exception(); }
    });
    struct Joiner {
        std::thread &t;
        ~Joiner() { if (t.joinable()) t.join(); }
    } power_joiner{power_thread};
    const cudaStream_t eng_st = st;
    st = upl.s;   // the instance stage below runs on the upload stream

    // host array (rows x B values) -> column-major staging buffer on the device
    auto stage_in = [&](double *dst, const double *host, int rows) {
        const size_t bytes = sizeof(double) * (size_t)rows * B;
        if (layout == 0) { h2d_large(dst, host, bytes, st); return; }
        h2d_large(S.d_raw, host, bytes, st);
        rowmajor_to_colmajor_kernel<<<dim3((rows + 31) / 32, (B + 31) / 32), 256, 0, st>>>(S.d_raw, dst, rows, B);
        S.launches++;
    };
    // per-instance scaling on the device -- reference build_batched_lp_device :792-885.  Inputs are staged straight from
    // the caller's (pageable) arrays by several host threads; no host copies, no host loops over B x (n + m) entries.
    const dim3 norm_grid(kNormChunks, B);
    const dim3 tgrid_n((n + 31) / 32, S.G), tgrid_m((m + 31) / 32, S.G);
    double *dn = S.d_norms;   // rows of Bpad: 0 norm_b_org, 1 b_scale, 2 norm_c_org, 3 c_scale, 4 norm_b, 5 norm_c
    // rows first: b_scale is needed by l and u
    stage_in(S.d_stage2, AL_in, m);
    stage_in(S.d_stage3, AU_in, m);
    batched_scale_rows_kernel<<<norm_grid, kBThreads, 0, st>>>(S.d_stage2, S.d_stage3, S.eng.row_norm, m, S.d_dd);
    batched_norm_finalize_kernel<<<(B + 127) / 128, 128, 0, st>>>(S.d_dd, 2, B, S.Bpad, dn, 0x3, bc ? 0 : 0x2);
    batched_div_rows_kernel<<<norm_grid, kBThreads, 0, st>>>(S.d_stage2, S.d_stage3, dn + S.Bpad, bc, m, S.d_dd);
    batched_norm_finalize_kernel<<<(B + 127) / 128, 128, 0, st>>>(S.d_dd, 1, B, S.Bpad, dn + 4 * (size_t)S.Bpad, 0, 0);
    to_group_layout_scaled_kernel<-1, false><<<tgrid_m, 256, 0, st>>>(S.d_stage2, S.AL, m, B, nullptr, nullptr, false);
    to_group_layout_scaled_kernel<1, false><<<tgrid_m, 256, 0, st>>>(S.d_stage3, S.AU, m, B, nullptr, nullptr, false);
    // cost
    stage_in(S.d_stage, C_in, n);
    batched_scale_cols_kernel<<<norm_grid, kBThreads, 0, st>>>(S.d_stage, S.eng.col_norm, n, S.d_dd);
    batched_norm_finalize_kernel<<<(B + 127) / 128, 128, 0, st>>>(S.d_dd, 2, B, S.Bpad, dn + 2 * (size_t)S.Bpad, 0x3, bc ? 0 : 0x2);
    batched_div_cols_kernel<<<norm_grid, kBThreads, 0, st>>>(S.d_stage, dn + 3 * (size_t)S.Bpad, bc, n, S.d_dd);
    batched_norm_finalize_kernel<<<(B + 127) / 128, 128, 0, st>>>(S.d_dd, 1, B, S.Bpad, dn + 5 * (size_t)S.Bpad, 0, 0);
    to_group_layout_scaled_kernel<0, false><<<tgrid_n, 256, 0, st>>>(S.d_stage, S.C, n, B, nullptr, nullptr, false);
    // bounds on x: l, u *= col_norm, /= b_scale, +-inf -> +-1e100
    stage_in(S.d_stage, l_in, n);
    to_group_layout_scaled_kernel<-1, true><<<tgrid_n, 256, 0, st>>>(S.d_stage, S.L, n, B, S.eng.col_norm, dn + S.Bpad, bc);
    stage_in(S.d_stage, u_in, n);
    to_group_layout_scaled_kernel<1, true><<<tgrid_n, 256, 0, st>>>(S.d_stage, S.U, n, B, S.eng.col_norm, dn + S.Bpad, bc);
    S.launches += 13;
    {
        std::vector<double> hn((size_t)6 * S.Bpad);
        HPR_CUDA_CHECK(cudaMemcpyAsync(hn.data(), dn, sizeof(double) * hn.size(), cudaMemcpyDeviceToHost, st));
        HPR_CUDA_CHECK(cudaStreamSynchronize(st));
        HPR_CUDA_CHECK(cudaGetLastError());
        auto rowk = [&](int r) { return std::vector<double>(hn.begin() + (size_t)r * S.Bpad, hn.begin() + (size_t)r * S.Bpad + B); };
        S.norm_b_org = rowk(0); S.b_scale = rowk(1); S.norm_c_org = rowk(2); S.c_scale = rowk(3); S.norm_b = rowk(4); S.norm_c = rowk(5);
    }
    S.obj_constants.assign(B, model->obj_constant);
    if (obj_constants) S.obj_constants.assign(obj_constants, obj_constants + B);
    stage_done("instance upload + scaling");
    st = eng_st;
    power_thread.join();
    if (power_error) std::rethrow_exception(power_error);

    RestartHost R;
    R.restart_flag.assign(B, 0); R.first_restart.assign(B, 1); R.inner.assign(B, 0); R.times.assign(B, 0);
    const double inf = std::numeric_limits<double>::infinity();
    R.last_gap.assign(B, inf); R.current_gap.assign(B, inf); R.save_gap.assign(B, inf); R.best_gap.assign(B, inf);
    R.sigma.assign(B, 1.0);
    for (int k = 0; k < B; ++k)
        if (S.norm_b[k] > 1.0e-8 && S.norm_c[k] > 1.0e-8) R.sigma[k] = S.norm_b[k] / S.norm_c[k];
    R.best_sigma = R.sigma;
    std::vector<double> sig_pad(S.Bpad, 1.0), bs_pad(S.Bpad, 1.0), cs_pad(S.Bpad, 1.0);
    std::vector<unsigned char> active(B, 1), act_pad(S.Bpad, 0), flag_pad(S.Bpad, 0);
    for (int k = 0; k < B; ++k) { sig_pad[k] = R.sigma[k]; bs_pad[k] = S.b_scale[k]; cs_pad[k] = S.c_scale[k]; act_pad[k] = 1; }
    HPR_CUDA_CHECK(cudaMemcpy(S.d_sigma, sig_pad.data(), sizeof(double) * S.Bpad, cudaMemcpyHostToDevice));
    HPR_CUDA_CHECK(cudaMemcpy(S.d_scale_b, bs_pad.data(), sizeof(double) * S.Bpad, cudaMemcpyHostToDevice));
    HPR_CUDA_CHECK(cudaMemcpy(S.d_scale_c, cs_pad.data(), sizeof(double) * S.Bpad, cudaMemcpyHostToDevice));
    HPR_CUDA_CHECK(cudaMemcpy(S.d_active, act_pad.data(), S.Bpad, cudaMemcpyHostToDevice));
    const double setup_time = now_s() - setup_start;
    stage_done("power iteration (rest) + scalars");

    // Result arrays (reference collect_results :887-935 allocates them after the loop): allocated now and touched by helper
    // threads while the GPU iterates, so the device-to-host copies at the end do not pay a page fault per 4 KB of fresh memory
    // (0.92 GB of results on configs[3]).
    struct OutArrays {
        double *x = nullptr, *y = nullptr, *z = nullptr;
        ~OutArrays() { std::free(x); std::free(y); std::free(z); }
    } outs;
    outs.x = static_cast<double *>(std::malloc(sizeof(double) * (size_t)n * B));
    outs.y = static_cast<double *>(std::malloc(sizeof(double) * (size_t)m * B));
    outs.z = static_cast<double *>(std::malloc(sizeof(double) * (size_t)n * B));
    if (!outs.x || !outs.y || !outs.z) throw std::bad_alloc();
    std::vector<std::thread> prefault;
    struct JoinAll {
        std::vector<std::thread> &ts;
        ~JoinAll() { for (auto &t : ts) if (t.joinable()) t.join(); }
    } prefault_joiner{prefault};
    if (sizeof(double) * ((size_t)2 * n + m) * B >= ((size_t)64 << 20)) {
        auto touch = [](double *p, size_t bytes, int part, int parts) {
            volatile char *c = reinterpret_cast<volatile char *>(p);
            const size_t lo = (bytes / parts * part) & ~(size_t)4095, hi = part + 1 == parts ? bytes : ((bytes / parts * (part + 1)) & ~(size_t)4095);
            for (size_t o = lo; o < hi; o += 4096) c[o] = 0;
        };
        for (int t = 0; t < 4; ++t)
            prefault.emplace_back([&, t]() {
                touch(outs.x, sizeof(double) * (size_t)n * B, t, 4);
                touch(outs.z, sizeof(double) * (size_t)n * B, t, 4);
                touch(outs.y, sizeof(double) * (size_t)m * B, t, 4);
            });
    }

    const double solve_start = now_s();
    ResidualHost res;
    res.primal_obj.assign(B, 0.0); res.dual_obj.assign(B, 0.0); res.err_Rp.assign(B, 0.0); res.err_Rd.assign(B, 0.0);
    res.rel_gap.assign(B, 0.0); res.kkt_error.assign(B, inf);
    std::vector<std::string> status(B, "CONTINUE");
    std::vector<int> final_iter(B, actual.max_iter);
    const int check_iter = std::max(actual.check_iter, 1);

    auto upload_active = [&]() {
        for (int k = 0; k < B; ++k) act_pad[k] = active[k];
        HPR_CUDA_CHECK(cudaMemcpyAsync(S.d_active, act_pad.data(), S.Bpad, cudaMemcpyHostToDevice, S.stream));
        HPR_CUDA_CHECK(cudaStreamSynchronize(S.stream));
    };

    int iter = 0;
    const char *final_status = nullptr;
    for (;;) {   // visited indices: multiples of check_iter and max_iter (reference loop :1017-1084)
        const bool periodic = (iter % check_iter) == 0;
        const double elapsed = now_s() - solve_start;
        if (periodic) {
            if (iter > 0) R.current_gap = S.weighted_norm(R.sigma);
            S.residuals(iter, &res);
            for (int k = 0; k < B; ++k)
                if (active[k] && res.kkt_error[k] <= actual.stop_tol) {   // `<=`: reference quirk #4
                    status[k] = "OPTIMAL";
                    final_iter[k] = iter;
                    active[k] = 0;
                }
            upload_active();
        }
        bool all_done = true;
        for (const std::string &s : status) all_done = all_done && (s != "CONTINUE");
        if (all_done) break;
        if (iter >= actual.max_iter || elapsed >= actual.time_limit) {
            final_status = elapsed >= actual.time_limit ? "TIME_LIMIT" : "ITER_LIMIT";
            for (int k = 0; k < B; ++k)
                if (status[k] == "CONTINUE") { status[k] = final_status; final_iter[k] = iter; active[k] = 0; }
            break;
        }
        std::fill(R.restart_flag.begin(), R.restart_flag.end(), 0);
        if (periodic) check_restart(&R, iter, check_iter, active);

        bool any = false;
        for (int f : R.restart_flag) any = any || (f >= 1 && f <= 3);
        if (any) {
            // update_sigma + do_restart (reference :702-762): movement norms for all, masked copies for the restarted
            for (int k = 0; k < B; ++k) flag_pad[k] = R.restart_flag[k] > 0 ? 1 : 0;
            HPR_CUDA_CHECK(cudaMemcpyAsync(S.d_flags, flag_pad.data(), S.Bpad, cudaMemcpyHostToDevice, S.stream));
            batched_restart_kernel<<<dim3(S.nbx_vec, S.G), kBThreads, 0, S.stream>>>(S.X_bar, S.lastX, S.X, n, S.Y_bar, S.lastY, S.Y, m,
                                                                                    S.d_flags, S.d_k, S.d_partials);
            batched_final_reduce_kernel<<<S.G, 32 * 2, 0, S.stream>>>(S.d_partials, S.nbx_vec, 2, S.Bpad, S.d_scal);
            S.launches += 2;
            S.fetch(2);
            const double sqrt_lambda = std::sqrt(S.lambda_max);
            for (int k = 0; k < B; ++k) {
                if (!active[k] || R.restart_flag[k] < 1) continue;
                const double pm = std::sqrt(S.hs(0, k)), dm = std::sqrt(S.hs(1, k));
                if (pm > 1.0e-16 && dm > 1.0e-16 && pm < 1.0e12 && dm < 1.0e12) {
                    const double ratio = (pm / dm) / sqrt_lambda;
                    const double fact = std::exp(-0.05 * (R.current_gap[k] / R.best_gap[k]));
                    const double temp1 = std::max(std::min(res.err_Rd[k], res.err_Rp[k]), std::min(res.rel_gap[k], R.current_gap[k]));
                    const double sigma_cand = std::exp(fact * std::log(ratio) + (1.0 - fact) * std::log(R.best_sigma[k]));
                    const double ratio_infeas = res.err_Rd[k] / res.err_Rp[k];
                    double kappa = 1.0;
                    if (temp1 > 9.0e-10) kappa = 1.0;
                    else if (temp1 > 5.0e-10) kappa = std::max(std::min(std::sqrt(ratio_infeas), 100.0), 1.0e-2);
                    else kappa = std::max(std::min(ratio_infeas, 100.0), 1.0e-2);
                    R.sigma[k] = kappa * sigma_cand;
                } else {
                    R.sigma[k] = 1.0;
                }
                R.times[k] += 1;
                R.inner[k] = 0;
                R.save_gap[k] = inf;
            }
            for (int k = 0; k < B; ++k) sig_pad[k] = R.sigma[k];
            HPR_CUDA_CHECK(cudaMemcpyAsync(S.d_sigma, sig_pad.data(), sizeof(double) * S.Bpad, cudaMemcpyHostToDevice, S.stream));
            HPR_CUDA_CHECK(cudaStreamSynchronize(S.stream));
        }

        // run to the next visited index; check iterations where the reference's to_check holds (:1067-1068)
        long long next = ((long long)iter / check_iter + 1) * check_iter;
        if ((long long)actual.max_iter > iter) next = std::min(next, (long long)actual.max_iter);
        const int stop = (int)std::min<long long>(next, INT32_MAX);
        for (int it = iter; it < stop; ++it) {
            const bool restarted = any && it == iter;
            const bool to_check = ((it + 1) % check_iter) == 0 || restarted || ((it + 1) % step_of(it + 1) == 0);
            S.iteration(to_check);
            if (restarted) {
                std::vector<double> lg = S.weighted_norm(R.sigma);
                for (int k = 0; k < B; ++k)
                    if (R.restart_flag[k] > 0) R.last_gap[k] = lg[k];
            }
        }
        for (int k = 0; k < B; ++k)
            if (active[k]) R.inner[k] += stop - iter;
        iter = stop;
    }
    HPR_CUDA_CHECK(cudaStreamSynchronize(S.stream));
    const double solve_time = now_s() - solve_start;
    stage_done("loop");

    // collect_results (reference :887-935): unscale on the device, one D2H per output array
    HPRLP_batched_results out;
    out.m = m; out.n = n; out.batch_size = B;
    for (auto &t : prefault) t.join();
    out.x = outs.x; out.y = outs.y; out.z = outs.z;   // still owned by `outs` until the copies below have succeeded
    out.primal_obj = static_cast<double *>(std::malloc(sizeof(double) * B));
    out.residuals = static_cast<double *>(std::malloc(sizeof(double) * B));
    out.gap = static_cast<double *>(std::malloc(sizeof(double) * B));
    out.iter = static_cast<int *>(std::malloc(sizeof(int) * B));
    out.status = static_cast<char *>(std::calloc(static_cast<size_t>(B) * 64, sizeof(char)));
    auto download = [&](const double *src, double *dst, int rows, bool divide, const double *norm, const double *scale) {
        if (divide) from_group_layout_kernel<true><<<dim3((rows + 31) / 32, S.G), 256, 0, S.stream>>>(src, S.d_stage, rows, B, norm, scale);
        else from_group_layout_kernel<false><<<dim3((rows + 31) / 32, S.G), 256, 0, S.stream>>>(src, S.d_stage, rows, B, norm, scale);
        d2h_large(dst, S.d_stage, sizeof(double) * (size_t)rows * B, S.stream);
    };
    download(S.X_bar, out.x, n, true, S.eng.col_norm, S.d_scale_b);
    download(S.Z_bar, out.z, n, false, S.eng.col_norm, S.d_scale_c);
    download(S.Y_bar, out.y, m, true, S.eng.row_norm, S.d_scale_c);
    for (int k = 0; k < B; ++k) {
        out.primal_obj[k] = res.primal_obj[k];
        out.residuals[k] = res.kkt_error[k];
        out.gap[k] = res.rel_gap[k];
        out.iter[k] = final_iter[k];
        std::strncpy(out.status + 64 * k, status[k].c_str(), 63);
    }
    out.setup_time = setup_time;
    out.solve_time = solve_time;
    out.power_time = power_time;
    out.time = setup_time + solve_time;
    stage_done("collect results");
    outs.x = outs.y = outs.z = nullptr;   // handed to the caller (free_batched_results)
    return out;
}

// Batch sharding over the GPUs of one node: instances are independent, A / scaling / lambda_max are
// replicated per GPU (deterministic: same cuRAND seed), no per-iteration collective.  One host thread per GPU.
// Known divergence (documented in DESIGN.md): the reference's shared lambda_max can be bumped by any instance of
// the batch (src/batched_solver.cu:642-646); here a bump stays local to the shard that saw it.
extern "C" HPRLP_batched_results hprlp_b200_solve_batched_multi(const LP_info_cpu *model, int batch_size, const HPRLP_FLOAT *C_in,
                                                                 const HPRLP_FLOAT *AL_in, const HPRLP_FLOAT *AU_in,
                                                                 const HPRLP_FLOAT *l_in, const HPRLP_FLOAT *u_in,
                                                                 const HPRLP_FLOAT *obj_constants, const HPRLP_parameters *param,
                                                                 int n_gpus) {
    if (!model || !model->A || batch_size <= 0 || !C_in || !AL_in || !AU_in || !l_in || !u_in) {
        return make_batched_error("ERROR", model ? model->m : 0, model ? model->n : 0, std::max(batch_size, 0));
    }
    // nothing is thrown across the C ABI: CUDA failures become "ERROR" status slots (reference :356-368 layout)
    return abi_guard<HPRLP_batched_results>("solve_batched", [&]() -> HPRLP_batched_results {
    int avail = 0;
    if (cudaGetDeviceCount(&avail) != cudaSuccess || avail < 1) throw std::runtime_error("solve_batched: no CUDA device");
    HPRLP_parameters def;
    const HPRLP_parameters base = param ? *param : def;
    int ng = std::max(1, std::min(std::min(n_gpus, avail), batch_size));
    if (ng == 1) return solve_batched_on_device(model, batch_size, C_in, AL_in, AU_in, l_in, u_in, obj_constants, &base);

    const int m = model->m, n = model->n, B = batch_size;
    std::vector<HPRLP_batched_results> parts(ng);
    std::vector<int> begin(ng + 1, 0);
    for (int d = 0; d < ng; ++d) begin[d + 1] = begin[d] + B / ng + (d < B % ng ? 1 : 0);   // contiguous shards
    std::vector<std::thread> workers;
    std::vector<std::string> errors(ng);
    for (int d = 0; d < ng; ++d) {
        workers.emplace_back([&, d]() {
            try {
                HPRLP_parameters p = base;
                p.device_number = base.device_number + d;
                const int k0 = begin[d], nb = begin[d + 1] - begin[d];
                parts[d] = solve_batched_on_device(model, nb, C_in + (size_t)k0 * n, AL_in + (size_t)k0 * m, AU_in + (size_t)k0 * m,
                                                   l_in + (size_t)k0 * n, u_in + (size_t)k0 * n,
                                                   obj_constants ? obj_constants + k0 : nullptr, &p);
            } catch (const std::exception &e) {
                errors[d] = e.what();
            }
        });
    }
    for (auto &w : workers) w.join();
    for (int d = 0; d < ng; ++d)
        if (!errors[d].empty()) throw std::runtime_error("solve_batched shard " + std::to_string(d) + ": " + errors[d]);
    HPRLP_batched_results out;
    out.m = m; out.n = n; out.batch_size = B;
    out.x = static_cast<double *>(std::malloc(sizeof(double) * (size_t)n * B));
    out.y = static_cast<double *>(std::malloc(sizeof(double) * (size_t)m * B));
    out.z = static_cast<double *>(std::malloc(sizeof(double) * (size_t)n * B));
    out.primal_obj = static_cast<double *>(std::malloc(sizeof(double) * B));
    out.residuals = static_cast<double *>(std::malloc(sizeof(double) * B));
    out.gap = static_cast<double *>(std::malloc(sizeof(double) * B));
    out.iter = static_cast<int *>(std::malloc(sizeof(int) * B));
    out.status = static_cast<char *>(std::calloc(static_cast<size_t>(B) * 64, sizeof(char)));
    for (int d = 0; d < ng; ++d) {
        const int k0 = begin[d], nb = begin[d + 1] - begin[d];
        HPRLP_batched_results &r = parts[d];
        std::memcpy(out.x + (size_t)k0 * n, r.x, sizeof(double) * (size_t)n * nb);
        std::memcpy(out.y + (size_t)k0 * m, r.y, sizeof(double) * (size_t)m * nb);
        std::memcpy(out.z + (size_t)k0 * n, r.z, sizeof(double) * (size_t)n * nb);
        std::memcpy(out.primal_obj + k0, r.primal_obj, sizeof(double) * nb);
        std::memcpy(out.residuals + k0, r.residuals, sizeof(double) * nb);
        std::memcpy(out.gap + k0, r.gap, sizeof(double) * nb);
        std::memcpy(out.iter + k0, r.iter, sizeof(int) * nb);
        std::memcpy(out.status + (size_t)64 * k0, r.status, (size_t)64 * nb);
        out.setup_time = std::max(out.setup_time, r.setup_time);
        out.solve_time = std::max(out.solve_time, r.solve_time);
        out.power_time = std::max(out.power_time, r.power_time);
        free_batched_results(&r);
    }
    out.time = out.setup_time + out.solve_time;
    return out;
    }, [&] { return make_batched_error("ERROR", model->m, model->n, batch_size); });
}

// The reference entry point (src/batched_solver.cu:939).  HPRLP_NUM_GPUS=N (environment) shards the batch over N
// GPUs starting at param->device_number; default 1 = the reference's single-GPU behaviour.
extern "C" HPRLP_batched_results solve_batched(const LP_info_cpu *model, int batch_size, const HPRLP_FLOAT *C_in,
                                                const HPRLP_FLOAT *AL_in, const HPRLP_FLOAT *AU_in, const HPRLP_FLOAT *l_in,
                                                const HPRLP_FLOAT *u_in, const HPRLP_FLOAT *obj_constants,
                                                const HPRLP_parameters *param) {
    int ng = 1;
    if (const char *e = getenv("HPRLP_NUM_GPUS")) ng = std::max(1, atoi(e));
    return hprlp_b200_solve_batched_multi(model, batch_size, C_in, AL_in, AU_in, l_in, u_in, obj_constants, param, ng);
}

// solve_batched for callers that hold the five dense inputs ROW-major, i.e. as C-ordered (n, B) / (m, B) arrays -- what the
// reference's Python API takes (bindings/python/src/hprlp_pybind.cpp:413-455 re-packs them element by element on the
// host before every call: SURVEY.md 8f rank 4).  Here the arrays are uploaded as they are and re-laid out on the device.
// layout: 0 = column-major (same as solve_batched), 1 = row-major.  Outputs are column-major as in solve_batched.
extern "C" HPRLP_batched_results hprlp_b200_solve_batched_layout(const LP_info_cpu *model, int batch_size, const HPRLP_FLOAT *C_in,
                                                                  const HPRLP_FLOAT *AL_in, const HPRLP_FLOAT *AU_in,
                                                                  const HPRLP_FLOAT *l_in, const HPRLP_FLOAT *u_in,
                                                                  const HPRLP_FLOAT *obj_constants, const HPRLP_parameters *param,
                                                                  int layout) {
    if (!model || !model->A || batch_size <= 0 || !C_in || !AL_in || !AU_in || !l_in || !u_in || layout < 0 || layout > 1) {
        return make_batched_error("ERROR", model ? model->m : 0, model ? model->n : 0, std::max(batch_size, 0));
    }
    return abi_guard<HPRLP_batched_results>("hprlp_b200_solve_batched_layout", [&]() -> HPRLP_batched_results {
        HPRLP_parameters def;
        const HPRLP_parameters base = param ? *param : def;
        return solve_batched_on_device(model, batch_size, C_in, AL_in, AU_in, l_in, u_in, obj_constants, &base, layout);
    }, [&] { return make_batched_error("ERROR", model->m, model->n, batch_size); });
}

extern "C" void free_batched_results(HPRLP_batched_results *results) {   // reference :1094-1105
    if (!results) return;
    std::free(results->x); std::free(results->y); std::free(results->z);
    std::free(results->primal_obj); std::free(results->residuals); std::free(results->gap);
    std::free(results->iter); std::free(results->status);
    *results = HPRLP_batched_results{};
}
