// engine.cu -- B200-native HPR-LP engine: device setup, scaling, power iteration, the
// Halpern/Peaceman-Rachford loop with restarts, residuals and solution collection.
// Host control flow restates the reference driver (src/HPRLP.cu:116-311, src/main_iterate.cu)
// exactly; every device operation is one of the hand-written kernels of kernels.cuh or the small
// vector kernels below.  No cuSPARSE / cuBLAS / cuSOLVER anywhere; cuRAND only produces the
// power-iteration start vector (setup), as the reference does (src/power_iteration.cu:44-52).
#include "engine.h"
#include "kernels.cuh"

#include <curand.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>


namespace hpr {

// ------------------------------------------------------------------------------------------------
// small vector kernels
// ------------------------------------------------------------------------------------------------
constexpr int kVecThreads = 256;
constexpr int kVecBlocks = 148 * 4;   // one wave of 4 CTAs per SM on B200

__global__ void final_reduce_kernel(const double *partials, int n_blocks, int ns, double *out) {
    // Each thread adds the blocks b = tid, tid + 1024, ... (all slots of a block in one 64-byte read), then a fixed xor tree
    // per warp and across warps: the summation order depends only on n_blocks, so every reduction is reproducible.
    __shared__ double sm[kMaxSlots][32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double v[kMaxSlots];
#pragma unroll
    for (int s = 0; s < kMaxSlots; ++s) v[s] = 0.0;
    for (int b = threadIdx.x; b < n_blocks; b += blockDim.x) {
        const double2 *p = reinterpret_cast<const double2 *>(partials + (size_t)b * kMaxSlots);
#pragma unroll
        for (int h = 0; h < kMaxSlots / 2; ++h) {
            if (2 * h < ns) {
                const double2 w = p[h];
                v[2 * h] += w.x;
                v[2 * h + 1] += w.y;
            }
        }
    }
#pragma unroll
    for (int s = 0; s < kMaxSlots; ++s) {
        if (s < ns) {
            double t = v[s];
            for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
            if (lane == 0) sm[s][wid] = t;
        }
    }
    __syncthreads();
    if (wid == 0) {
        for (int s = 0; s < ns; ++s) {
            double t = (lane < (int)(blockDim.x >> 5)) ? sm[s][lane] : 0.0;
            for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
            if (lane == 0) out[s] = t;
        }
    }
}

template <int NS>
__device__ __forceinline__ void vec_block_store(double (&acc)[NS], double *partials) {
    __shared__ double scratch[kMaxSlots * (kVecThreads / 32)];
    block_reduce_store<NS>(acc, partials, scratch);
}

// |b|^2 with b_i = max(|AL_i|,|AU_i|), +-inf -> 0 (reference conceptual_b_kernel + cublasDnrm2,
// HPR_cuda_kernels.cu:34-43, src/scaling.cu:113-116); slot 1: |c|^2.
__global__ void __launch_bounds__(kVecThreads) norm_bc_kernel(const double *AL, const double *AU, int m, const double *c, int n,
                                                             double *partials) {
    double t[2] = {0.0, 0.0};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
        double a = AL[i], b = AU[i];
        a = isinf(a) ? 0.0 : a;
        b = isinf(b) ? 0.0 : b;
        const double v = fmax(fabs(a), fabs(b));
        t[0] += v * v;
    }
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) t[1] += c[j] * c[j];
    vec_block_store<2>(t, partials);
}

__global__ void __launch_bounds__(kVecThreads) sumsq_kernel(const double *v, int len, double *partials) {
    double t[1] = {0.0};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < len; i += gridDim.x * blockDim.x) t[0] += v[i] * v[i];
    vec_block_store<1>(t, partials);
}

// restart: movement norms |x_bar-x0|^2, |y_bar-y0|^2 (reference update_sigma axpby+nrm2,
// src/main_iterate.cu:370-376) fused with do_restart's four copies (:312-322).
__global__ void __launch_bounds__(kVecThreads) restart_kernel(const double *x_bar, double *x0, double *x, int j0, int j1, const double *y_bar,
                                                             double *y0, double *y, int m, double *partials) {
    double t[2] = {0.0, 0.0};
    for (int j = j0 + blockIdx.x * blockDim.x + threadIdx.x; j < j1; j += gridDim.x * blockDim.x) {
        const double xb = x_bar[j];
        const double d = xb - x0[j];
        t[0] += d * d;
        x0[j] = xb;
        x[j] = xb;
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
        const double yb = y_bar[i];
        const double d = yb - y0[i];
        t[1] += d * d;
        y0[i] = yb;
        y[i] = yb;
    }
    vec_block_store<2>(t, partials);
}

// q = z / sqrt(<z,z> + eps)  (reference src/power_iteration.cu:62-70), <z,z> read from a device scalar.
__global__ void power_normalize_kernel(const double *z, double *q, const double *zz, int m) {
    const double invn = 1.0 / sqrt(zz[0] + 2.220446049250313e-16);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) q[i] = invn * z[i];
}

// |z - lambda q|^2 with lambda = <q,z> read from a device scalar (reference :85-93).
__global__ void __launch_bounds__(kVecThreads) power_error_kernel(const double *z, const double *q, const double *lambda, int m, double *partials) {
    const double lam = lambda[0];
    double t[1] = {0.0};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
        const double r = fma(-lam, q[i], z[i]);
        t[0] += r * r;
    }
    vec_block_store<1>(t, partials);
}

__global__ void set_params_kernel(double *params, double sigma, double lambda_max) {
    const double f = lambda_max * sigma;
    params[0] = sigma;
    params[1] = f;
    params[2] = 1.0 / f;
    params[3] = 1.0 / sigma;
}

__global__ void fill_kernel(double *v, int len, double value) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < len; i += gridDim.x * blockDim.x) v[i] = value;
}
__global__ void add_scalar_kernel(double *v, int len, double value) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < len; i += gridDim.x * blockDim.x) v[i] += value;
}
__global__ void scal_kernel(double *v, int len, double alpha) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < len; i += gridDim.x * blockDim.x) v[i] = alpha * v[i];
}
__global__ void exp_clamp_kernel(double *v, int len) {   // reference src/scaling.cu:33-38
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < len; i += gridDim.x * blockDim.x)
        v[i] = fmin(fmax(exp(v[i]), 1e-30), 1e30);
}
// Row-side vector updates of one scaling stage.  CR multiplies (norm /= t, AL,AU *= t;
// src/scaling.cu:69-80), Ruiz/PC divide (norm *= t, AL,AU /= t; :129-133).
template <bool CR>
__global__ void scale_row_vectors_kernel(const double *t, double *norm, double *AL, double *AU, int m) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
        const double f = t[i];
        if (CR) { norm[i] = norm[i] / f; AL[i] = AL[i] * f; AU[i] = AU[i] * f; }
        else    { norm[i] = norm[i] * f; AL[i] = AL[i] / f; AU[i] = AU[i] / f; }
    }
}
// Column side: CR  norm /= t, c *= t, l,u /= t ; Ruiz/PC  norm *= t, c /= t, l,u *= t.
template <bool CR>
__global__ void scale_col_vectors_kernel(const double *t, double *norm, double *c, double *l, double *u, int n) {
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const double f = t[j];
        if (CR) { norm[j] = norm[j] / f; c[j] = c[j] * f; l[j] = l[j] / f; u[j] = u[j] / f; }
        else    { norm[j] = norm[j] * f; c[j] = c[j] / f; l[j] = l[j] * f; u[j] = u[j] * f; }
    }
}
// x = b_scale (x_bar / col_norm), z = c_scale (z_bar * col_norm), y = c_scale (y_bar / row_norm)
// (reference collect_solution, src/utils.cu:172-189: elementwise op, then a separate scale).
__global__ void unscale_kernel(const double *x_bar, const double *z_bar, const double *col_norm, double *xo, double *zo, int n,
                               const double *y_bar, const double *row_norm, double *yo, int m, double bs, double cs) {
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const double cn = col_norm[j];
        double a = x_bar[j] / cn;
        double b = z_bar[j] * cn;
        xo[j] = bs * a;
        zo[j] = cs * b;
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
        double a = y_bar[i] / row_norm[i];
        yo[i] = cs * a;
    }
}

// ---- row-partitioned mode: unfused x-side kernels on the x-block [j0, j1) this GPU owns --------------------
// x-update from the reduce-scattered w = (A^T y)_{J_p} (same arithmetic as XPhaseOp::row)
template <bool CHECK>
__global__ void __launch_bounds__(kVecThreads) x_update_kernel(const double *w, double *x, double *x_hat, const double *c, const double *l,
                                                              const double *u, const double *x0, double *x_bar, double *z_bar,
                                                              double *x_tmp, const double *params, const int *kx, int *ky, int j0, int j1) {
    const double sigma = params[0];
    const int k = *kx;
    const double f1 = 1.0 / (k + 2.0), f2 = 1.0 - f1;
    if (blockIdx.x == 0 && threadIdx.x == 0) *ky = k;
    for (int j = j0 + blockIdx.x * blockDim.x + threadIdx.x; j < j1; j += gridDim.x * blockDim.x) {
        const double xi = x[j];
        const double zt = fma(sigma, w[j] - c[j], xi);
        const double xb = fmin(u[j], fmax(l[j], zt));
        const double xh = 2.0 * xb - xi;
        x[j] = fma(f2, xh, f1 * x0[j]);
        x_hat[j] = xh;
        if (CHECK) { x_bar[j] = xb; z_bar[j] = (xb - zt) / sigma; x_tmp[j] = xb - xh; }
    }
}
// ---- the same exchange + x-update as ONE kernel over NVLink peer memory (collective.h, PeerExchange) ------------------
struct PeerPtrs {
    double *xhat[kMaxPeers];
    unsigned long long *flags[kMaxPeers];
};
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// Polls a flag of THIS GPU's memory that a peer GPU writes.  Different GPUs run concurrently, so the wait is finite; the
// bound turns a protocol bug or a dead peer into a trapped context (-> "ERROR" result) instead of a hung GPU.
__device__ __forceinline__ void wait_epoch(const unsigned long long *flag, unsigned long long epoch) {
    unsigned spins = 0;
    while (ld_acquire_sys(flag) < epoch) {
        __nanosleep(100);
        if (++spins > (1u << 27)) __trap();
    }
}
// "my partial w (slot 0) / my x_hat stores (slot 1) of this epoch are complete": one release store into every rank's flags
__global__ void exchange_signal_kernel(PeerPtrs pp, int P, int rank, int slot, unsigned long long epoch) {
    if ((int)threadIdx.x < P) {
        __threadfence_system();
        st_release_sys(pp.flags[threadIdx.x] + slot * kMaxPeers + rank, epoch);
    }
}
__global__ void exchange_wait_kernel(const unsigned long long *flags, int P, int slot, unsigned long long epoch) {
    if ((int)threadIdx.x < P) wait_epoch(flags + slot * kMaxPeers + threadIdx.x, epoch);
}
// reduce (of the scattered partials) + x-update + all-gather of one HPR iteration for the x-block [j0, j1) this GPU owns:
//   wait until every rank's A_q^T y_q pass has delivered its slot (the passes push their rows into `recv`, SpmvPushOp);
//   w_j = sum over the P local slots in rank order (the result does not depend on who finished first);  x-update (same
//   arithmetic as XPhaseOp::row);  x_hat_j stored into EVERY rank's x_hat buffer (P2P stores over NVLink);  the last CTA
//   tells every rank that this block of x_hat is in place.
// PP = number of ranks (compile time: unrolled) or 0 (any rank count).  16-byte accesses; j0 is even.
// MODE 0: the x-update above.  The other modes reuse the kernel's wait / push / signal frame for the remaining exchanges of
// the partitioned mode (the x_hat buffers are free scratch whenever they run):
//   1  sum of the slots -> every rank's x_hat           (all-reduce of pushed partials: A^T q of the power iteration)
//   2  sum of the slots -> my own x_hat block only      (reduce-scatter alone: (A^T y_bar) on the owned block, dual residual)
//   3  src block [j0, j1) -> every rank's x_hat         (all-gather alone: x_bar / x_bar - x_hat for the primal residual passes)
template <bool CHECK, int PP, int MODE = 0>
__global__ void __launch_bounds__(kVecThreads, 4) fused_exchange_x_kernel(PeerPtrs pp, const double *recv, size_t xblock, int P, int rank,
                                                                          unsigned long long epoch, unsigned *done, double *x,
                                                                          const double *c, const double *l, const double *u, const double *x0,
                                                                          double *x_bar, double *z_bar, double *x_tmp, const double *params,
                                                                          const int *kx, int *ky, int j0, int j1) {
    if ((int)threadIdx.x < P) wait_epoch(pp.flags[rank] + threadIdx.x, epoch);
    __syncthreads();
    constexpr bool PLAIN = MODE != 0;
    const double sigma = PLAIN ? 1.0 : params[0];
    const int k = PLAIN ? 0 : *kx;
    const double f1 = 1.0 / (k + 2.0), f2 = 1.0 - f1;
    if (!PLAIN && blockIdx.x == 0 && threadIdx.x == 0) *ky = k;
    auto update = [&](int j, double w, double xi, double cj, double lj, double uj, double x0j, double &xn, double &xh) {
        if (PLAIN) { xh = w; xn = 0.0; return; }
        const double zt = fma(sigma, w - cj, xi);
        const double xb = fmin(uj, fmax(lj, zt));
        xh = 2.0 * xb - xi;
        xn = fma(f2, xh, f1 * x0j);
        if (CHECK) { x_bar[j] = xb; z_bar[j] = (xb - zt) / sigma; x_tmp[j] = xb - xh; }
    };
    const int npairs = (j1 - j0) >> 1;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < npairs; t += gridDim.x * blockDim.x) {
        const int j = j0 + 2 * t;
        const double *rw = recv + 2 * t;   // slot q of this pair at rw + q * xblock
        double2 w = make_double2(0.0, 0.0);
        if (MODE == 3) {
            w = *reinterpret_cast<const double2 *>(recv + 2 * t);   // recv = source vector at j0
        } else if (PP > 0) {
            double2 wv[PP > 0 ? PP : 1];
#pragma unroll
            for (int q = 0; q < PP; ++q) wv[q] = __ldcs(reinterpret_cast<const double2 *>(rw + (size_t)q * xblock));
#pragma unroll
            for (int q = 0; q < PP; ++q) { w.x += wv[q].x; w.y += wv[q].y; }
        } else {
            for (int q = 0; q < P; ++q) { const double2 v = __ldcs(reinterpret_cast<const double2 *>(rw + (size_t)q * xblock)); w.x += v.x; w.y += v.y; }
        }
        double2 xi = make_double2(0.0, 0.0), cj = xi, lj = xi, uj = xi, xa = xi;
        if (!PLAIN) {
            xi = *reinterpret_cast<const double2 *>(x + j); cj = *reinterpret_cast<const double2 *>(c + j);
            lj = *reinterpret_cast<const double2 *>(l + j); uj = *reinterpret_cast<const double2 *>(u + j);
            xa = *reinterpret_cast<const double2 *>(x0 + j);
        }
        double2 xn, xh;
        update(j, w.x, xi.x, cj.x, lj.x, uj.x, xa.x, xn.x, xh.x);
        update(j + 1, w.y, xi.y, cj.y, lj.y, uj.y, xa.y, xn.y, xh.y);
        if (!PLAIN) *reinterpret_cast<double2 *>(x + j) = xn;
        if (MODE == 2) {
            *reinterpret_cast<double2 *>(pp.xhat[rank] + j) = xh;
        } else if (PP > 0) {
#pragma unroll
            for (int q = 0; q < PP; ++q) *reinterpret_cast<double2 *>(pp.xhat[q] + j) = xh;
        } else {
            for (int q = 0; q < P; ++q) *reinterpret_cast<double2 *>(pp.xhat[q] + j) = xh;
        }
    }
    if (((j1 - j0) & 1) && blockIdx.x == 0 && threadIdx.x == 0) {   // odd block length: the last entry
        const int j = j1 - 1;
        double w = 0.0;
        if (MODE == 3) w = recv[j - j0];
        else for (int q = 0; q < P; ++q) w += recv[(size_t)q * xblock + (j - j0)];
        double xn, xh;
        if (PLAIN) update(j, w, 0.0, 0.0, 0.0, 0.0, 0.0, xn, xh);
        else { update(j, w, x[j], c[j], l[j], u[j], x0[j], xn, xh); x[j] = xn; }
        if (MODE == 2) pp.xhat[rank][j] = xh;
        else for (int q = 0; q < P; ++q) pp.xhat[q][j] = xh;
    }
    if (MODE == 2) return;    // nothing was stored to a peer: nobody waits for this rank
    __threadfence_system();   // this thread's peer stores are performed before the CTA is counted as done
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(done, 1u);
        if (prev == gridDim.x - 1) {
            *done = 0u;
            __threadfence_system();
            for (int q = 0; q < P; ++q) st_release_sys(pp.flags[q] + kMaxPeers + rank, epoch);
        }
    }
}

// dual residual terms from the reduce-scattered w = (A^T y_bar)_{J_p} (same slots as ResidualDualOp; sums over J_p)
template <bool GAP, bool ITER0>
__global__ void __launch_bounds__(kVecThreads) residual_dual_kernel(const double *w, const double *c, const double *z_bar, const double *x_bar,
                                                                   const double *x_tmp, const double *col_norm, const double *l,
                                                                   const double *u, int j0, int j1, double *partials) {
    double t[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    for (int j = j0 + blockIdx.x * blockDim.x + threadIdx.x; j < j1; j += gridDim.x * blockDim.x) {
        const double cj = c[j], zb = z_bar[j], xb = x_bar[j], cn = col_norm[j];
        const double rd = (cj - w[j] - zb) * cn;
        t[0] += rd * rd; t[1] += cj * xb; t[2] += xb * zb;
        if (GAP) { const double dx = x_tmp[j]; t[3] += dx * dx; }
        if (ITER0) {
            const double lj = l[j], uj = u[j];
            const double viol = (xb < lj) ? (lj - xb) : ((xb > uj) ? (xb - uj) : 0.0);
            const double q = viol / cn;
            t[4] += q * q;
        }
    }
    vec_block_store<5>(t, partials);
}
// column statistics after the cross-GPU reduction: sqrt + clamp (Ruiz / Pock-Chambolle), mean (Curtis-Reid)
__global__ void sqrt_clamp_kernel(double *v, int len) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < len; i += gridDim.x * blockDim.x) {
        double r = sqrt(v[i]);
        if (r < 1e-15) r = 1.0;
        v[i] = r;
    }
}
__global__ void mean_kernel(double *sum, const double *cnt, int len) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < len; i += gridDim.x * blockDim.x)
        sum[i] = cnt[i] > 0.0 ? sum[i] / cnt[i] : 0.0;
}

static inline int vec_grid(int len) {
    int g = (len + kVecThreads - 1) / kVecThreads;
    return std::max(1, std::min(g, kVecBlocks));
}

// ------------------------------------------------------------------------------------------------
// launch helpers
// ------------------------------------------------------------------------------------------------
// per-CTA partial blocks written by one streamed pass (a banded pass reduces in its last band only)
static int part_blocks(const DevCsr &M) { return M.bands.empty() ? M.n_items : M.bands.back().n_items; }

static CsrView<int> view_of(const DevCsr &M) {
    CsrView<int> v;
    v.rows = M.rows; v.nnz = M.nnz; v.rowPtr = M.rowPtr; v.col = M.col; v.val = M.val;
    v.item_row = M.item_row; v.n_items = M.n_items;
    v.head_part = M.head_part; v.tail_part = M.tail_part; v.ticket = M.ticket;
    v.issued_host = const_cast<unsigned long long *>(&M.tickets_issued);
    v.carry_in = nullptr; v.carry_out = nullptr; v.chunk_offset = 0;
    return v;
}

// tickets after which the (32-bit) chunk counter of a matrix is re-zeroed between launches; HPRLP_TICKET_WRAP: tests
static unsigned long long ticket_wrap() {
    static const unsigned long long w = getenv("HPRLP_TICKET_WRAP") ? strtoull(getenv("HPRLP_TICKET_WRAP"), nullptr, 10) : 0xC0000000ull;
    return w;
}

template <class Op, int G>
static void launch_one(const CsrView<int> &v, const Op &op, cudaStream_t st) {
    constexpr size_t bytes = stream_smem_bytes<Op>();
    static std::atomic<unsigned> configured{0};   // per instantiation, one bit per device (function attributes are per device)
    int dev = 0;
    cudaGetDevice(&dev);
    if (!(configured.load() & (1u << dev))) {
        HPR_CUDA_CHECK(cudaFuncSetAttribute(csr_stream_kernel<Op, G, int>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        if (const char *e = getenv("HPRLP_CARVEOUT"))   // tuning hook: shared-memory carve-out in percent
            HPR_CUDA_CHECK(cudaFuncSetAttribute(csr_stream_kernel<Op, G, int>, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(e)));
        configured.fetch_or(1u << dev);
    }
    // ticket counter: a multiple of n_items between launches; zeroed here (stream-ordered) long before 2^32
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cap);
    if (cap == cudaStreamCaptureStatusNone) {   // (captured launches are accounted for per replay in run_normal)
        if (*v.issued_host + (unsigned long long)v.n_items > ticket_wrap()) {
            HPR_CUDA_CHECK(cudaMemsetAsync(v.ticket, 0, sizeof(unsigned), st));
            *v.issued_host = 0;
        }
        *v.issued_host += (unsigned long long)v.n_items;
    }
    // programmatic stream serialization: the grid may become resident while the previous kernel of the stream drains
    // (kernels.cuh: its CTAs wait at griddepcontrol.wait before touching anything an earlier kernel wrote)
    static const bool pdl = getenv("HPRLP_NO_PDL") == nullptr;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)v.n_items); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = bytes; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    static const bool pdl_in_graphs = getenv("HPRLP_NO_PDL_GRAPH") == nullptr;   // captured launches become programmatic edges
    cfg.attrs = attr; cfg.numAttrs = (pdl && (cap == cudaStreamCaptureStatusNone || pdl_in_graphs)) ? 1 : 0;
    HPR_CUDA_CHECK(cudaLaunchKernelEx(&cfg, csr_stream_kernel<Op, G, int>, v, op));
}

template <class Op>
static void dispatch_hot(const CsrView<int> &v, int G, const Op &op, cudaStream_t st) {
    switch (G) {
        case 1:  launch_one<Op, 1>(v, op, st); break;
        case 2:  launch_one<Op, 2>(v, op, st); break;
        case 4:  launch_one<Op, 4>(v, op, st); break;
        case 8:  launch_one<Op, 8>(v, op, st); break;
        case 16: launch_one<Op, 16>(v, op, st); break;
        default: launch_one<Op, 32>(v, op, st); break;
    }
}
// setup / check-iteration passes: fewer instantiations
template <class Op>
static void dispatch_cold(const CsrView<int> &v, int G, const Op &op, cudaStream_t st) {
    if (G <= 2)      launch_one<Op, 1>(v, op, st);
    else if (G <= 8) launch_one<Op, 4>(v, op, st);
    else             launch_one<Op, 16>(v, op, st);
}
// A pass over a column-banded matrix: one launch per band, row sums carried from band to band, the op's epilogue (and its
// reductions) only in the last one.  The gathered slice of every launch fits the L2.
template <class Op, bool HOT>
static void launch_banded(const DevCsr &M, const Op &op, cudaStream_t st) {
    static_assert(Op::NV == 1 && !Op::kMax, "banded passes carry one sum per row");
    const int nb = (int)M.bands.size();
    for (int b = 0; b < nb; ++b) {
        CsrView<int> v = view_of(M.bands[b]);
        v.carry_in = b > 0 ? M.carry : nullptr;
        v.carry_out = b + 1 < nb ? M.carry : nullptr;
        if (HOT) dispatch_hot(v, M.bands[b].G, op, st); else dispatch_cold(v, M.bands[b].G, op, st);
    }
}
template <class Op>
static void launch_stream_hot(const DevCsr &M, const Op &op, cudaStream_t st) {
    if constexpr (Op::NV == 1 && !Op::kMax) {
        if (!M.bands.empty()) { launch_banded<Op, true>(M, op, st); return; }
    }
    dispatch_hot(view_of(M), M.G, op, st);
}
template <class Op>
static void launch_stream(const DevCsr &M, const Op &op, cudaStream_t st) {
    if constexpr (Op::NV == 1 && !Op::kMax) {
        if (!M.bands.empty()) { launch_banded<Op, false>(M, op, st); return; }
    }
    dispatch_cold(view_of(M), M.G, op, st);
}

// out = M g over the hot dispatch (power iteration, partitioned x-side pass); tex = g as a texture, or 0
template <bool DOTS>
static void launch_spmv_hot(const DevCsr &M, const double *g, cudaTextureObject_t tex, double *out, const double *q, double *partials,
                            cudaStream_t st) {
    if (tex) { SpmvOp<DOTS, true> o; o.g = g; o.tex = tex; o.out = out; o.q = q; o.partials = partials; launch_stream_hot(M, o, st); }
    else     { SpmvOp<DOTS, false> o; o.g = g; o.tex = 0; o.out = out; o.q = q; o.partials = partials; launch_stream_hot(M, o, st); }
}

// One device arena per engine: a single cudaMalloc + one zero-fill instead of ~45 cudaMalloc/cudaMemset/cudaFree
// pairs (cudaMalloc/cudaFree of multi-GB buffers are synchronous and cost tens of ms each on the e2e path).
struct Arena { char *base = nullptr; size_t size = 0, off = 0; };
static thread_local Arena *g_arena = nullptr;   // set while an engine carves its buffers
static size_t arena_round(size_t bytes) { return (bytes + 511) / 512 * 512; }   // 512 B: texture base alignment

template <typename T>
static T *dalloc(size_t count) {
    const size_t bytes = arena_round(std::max<size_t>(count, 1) * sizeof(T));
    if (g_arena) {
        if (g_arena->off + bytes > g_arena->size) throw std::runtime_error("device arena exhausted");
        T *p = reinterpret_cast<T *>(g_arena->base + g_arena->off);   // arena memory is already zero
        g_arena->off += bytes;
        return p;
    }
    T *p = nullptr;
    HPR_CUDA_CHECK(cudaMalloc(&p, bytes));
    HPR_CUDA_CHECK(cudaMemset(p, 0, bytes));
    // cudaMemset runs on the legacy default stream and is asynchronous for device memory; the engine's streams are
    // non-blocking, so without this barrier a later kernel could be overtaken by the zero-fill.
    HPR_CUDA_CHECK(cudaDeviceSynchronize());
    return p;
}
static void dfree(void *p) { if (p) cudaFree(p); }

// Small pinned host blocks (residual scalars, sigma parameters) are recycled across solves for the life of the process:
// cudaFreeHost synchronises the device and was measured at 0.02 - 2.5 s per call on B200 after multi-GB solves
// (HPRLP_TIMING=1), more than the whole C2 solve.  Blocks are portable (any device), 32 doubles each.
namespace {
std::mutex g_pinned_mu;
std::vector<double *> g_pinned_free;
}
double *pinned_block_acquire() {
    {
        std::lock_guard<std::mutex> lk(g_pinned_mu);
        if (!g_pinned_free.empty()) { double *p = g_pinned_free.back(); g_pinned_free.pop_back(); return p; }
    }
    double *p = nullptr;
    HPR_CUDA_CHECK(cudaHostAlloc(reinterpret_cast<void **>(&p), 32 * sizeof(double), cudaHostAllocPortable));
    return p;
}
void pinned_block_release(double *p) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_pinned_mu);
    g_pinned_free.push_back(p);
}

// The engines' private stream-ordered pool, one per device, created on first use.
namespace {
constexpr int kMaxDevices = 64;
std::mutex g_pool_mu;
cudaMemPool_t g_pools[kMaxDevices] = {};
}
static cudaMemPool_t engine_pool(int device) {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    if (device < 0 || device >= kMaxDevices) throw std::runtime_error("device number out of range");
    if (!g_pools[device]) {
        cudaMemPoolProps props{};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        HPR_CUDA_CHECK(cudaMemPoolCreate(&g_pools[device], &props));
        unsigned long long keep = 8192ULL << 20;
        if (const char *e = getenv("HPRLP_POOL_RETAIN_MB")) keep = strtoull(e, nullptr, 10) << 20;
        HPR_CUDA_CHECK(cudaMemPoolSetAttribute(g_pools[device], cudaMemPoolAttrReleaseThreshold, &keep));
    }
    return g_pools[device];
}
// One cuRAND generator per device, kept across solves: creating and destroying one costs a device allocation, a free (a
// device-wide synchronisation) and 1-250 ms of wall time per solve on the B200 boxes (HPRLP_TIMING, "power start vector").
// Re-seeding and rewinding it reproduces the sequence of a fresh generator.
struct CachedGenerator {
    std::mutex mu;
    curandGenerator_t gen = nullptr;
};
static CachedGenerator g_gens[kMaxDevices];

// N(0,1) doubles of a fresh XORWOW generator with seed 1 (reference src/preprocess.cu:153-156), count even
static void draw_power_start(int device, double *out, size_t count, cudaStream_t stream) {
    if (device < 0 || device >= kMaxDevices) throw std::runtime_error("device number out of range");
    CachedGenerator &c = g_gens[device];
    std::lock_guard<std::mutex> lk(c.mu);
    if (!c.gen && curandCreateGenerator(&c.gen, CURAND_RNG_PSEUDO_DEFAULT) != CURAND_STATUS_SUCCESS) {
        c.gen = nullptr;
        throw std::runtime_error("curandCreateGenerator failed");
    }
    bool ok = curandSetStream(c.gen, stream) == CURAND_STATUS_SUCCESS;
    ok = ok && curandSetPseudoRandomGeneratorSeed(c.gen, 1ULL) == CURAND_STATUS_SUCCESS;
    ok = ok && curandSetGeneratorOffset(c.gen, 0ULL) == CURAND_STATUS_SUCCESS;
    ok = ok && curandGenerateNormalDouble(c.gen, out, count, 0.0, 1.0) == CURAND_STATUS_SUCCESS;
    const cudaError_t e = cudaStreamSynchronize(stream);   // the generator must not be re-seeded under a running draw
    if (!ok || e != cudaSuccess) {
        curandDestroyGenerator(c.gen);
        c.gen = nullptr;
        throw std::runtime_error("cuRAND draw of the power-iteration start vector failed");
    }
}

void release_cached_device_memory() {
    int prev = 0;
    cudaGetDevice(&prev);
    for (int d = 0; d < kMaxDevices; ++d) {
        std::lock_guard<std::mutex> lk(g_gens[d].mu);
        if (g_gens[d].gen) {
            cudaSetDevice(d);
            curandDestroyGenerator(g_gens[d].gen);
            g_gens[d].gen = nullptr;
        }
    }
    cudaSetDevice(prev);
    std::lock_guard<std::mutex> lk(g_pool_mu);
    for (cudaMemPool_t p : g_pools)
        if (p) HPR_CUDA_CHECK(cudaMemPoolTrimTo(p, 0));
}

static double now_seconds() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// reference src/utils.cu:100-102
static int step_of(int iter) {
    return std::max(10, static_cast<int>(pow(10, floor(log10((double)iter))) / 10));
}

// ------------------------------------------------------------------------------------------------
// setup
// ------------------------------------------------------------------------------------------------
void csr_transpose_host(int rows, int cols, int nnz, const int *rp, const int *ci, const double *v,
                        int *trp, int *tci, double *tv) {
    // Stable counting sort by column: entries of a transposed row keep the original row order,
    // i.e. the same entry order as the reference's CSR_transpose_host (src/utils.cu:203-232).
    std::vector<int> cursor((size_t)cols + 1, 0);
    for (int k = 0; k < nnz; ++k) cursor[(size_t)ci[k] + 1]++;
    for (int j = 0; j < cols; ++j) cursor[j + 1] += cursor[j];
    for (int j = 0; j <= cols; ++j) trp[j] = cursor[j];
    for (int i = 0; i < rows; ++i)
        for (int k = rp[i]; k < rp[i + 1]; ++k) {
            const int pos = cursor[ci[k]]++;
            tci[pos] = i;
            tv[pos] = v[k];
        }
}

static void alloc_matrix(DevCsr &M, int rows, int cols, long long nnz) {
    M.rows = rows; M.cols = cols; M.nnz = nnz;
    M.n_items = (int)((nnz + kChunk - 1) / kChunk);
    if (M.n_items < 1) M.n_items = 1;
    const size_t padded = (size_t)M.n_items * kChunk;
    M.rowPtr = dalloc<int>((size_t)rows + 1);
    M.col = dalloc<int>(padded);
    M.val = dalloc<double>(padded);
    const size_t witems = (size_t)M.n_items * kWarps;   // warp items (n_items = CTAs)
    M.item_row = dalloc<int>(witems + 1);
    M.head_part = dalloc<PartSlot>(witems * 2);   // all-ones = "not published" (set in finish_matrix); consumers re-arm what they read
    M.tail_part = dalloc<PartSlot>(witems * 2);
    M.ticket = dalloc<unsigned>(1);
}
static void free_matrix(DevCsr &M) {
    dfree(M.rowPtr); dfree(M.col); dfree(M.val); dfree(M.item_row);
    dfree(M.head_part); dfree(M.tail_part); dfree(M.ticket);
    M = DevCsr();
}

static int pick_lanes(double mean_len, double len_cv, const char *env_name) {
    if (const char *e = getenv(env_name)) {
        const int g = atoi(e);
        if (g == 1 || g == 2 || g == 4 || g == 8 || g == 16 || g == 32) return g;
    }
    // Lanes per row from the measured row-length statistics of this matrix.  Every extra lane costs shuffle
    // wavefronts on the L1 data stage (the binding unit), so few lanes win: B200 sweep on C2/C3
    // (gpurun_out_14/15): mean 10 and 20 -> 1 lane best; mean 50 and 100 -> 8 lanes best (4 and 16 within 3 %).
    // With one lane per row the batch of 32 rows takes as long as its longest row, so rows of uneven length want lanes
    // even when they are short on average (profiles/r2_lanes_vs_structure.md: mean 20, std/mean 0.63 -> 4 lanes 19 % faster,
    // 0.45 -> 7 %, 0.32 -> 3 %; Poisson lengths of a uniformly random matrix, 0.22 -> 1 lane 5 % faster).
    if (mean_len < 8.0) return 1;
    if (mean_len < 28.0) return len_cv >= 0.4 ? 4 : 1;
    if (mean_len < 40.0) return 4;
    return 8;
}

// sum over rows of len^2 (exact, integer): the spread of the row lengths decides the lanes per row (pick_lanes)
__global__ void row_len_sq_kernel(const int *rowPtr, int rows, unsigned long long *out) {
    unsigned long long acc = 0;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += gridDim.x * blockDim.x) {
        const unsigned long long len = (unsigned long long)(rowPtr[r + 1] - rowPtr[r]);
        acc += len * len;
    }
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

void Engine::finish_matrix(DevCsr &M) {
    const int threads = 256;
    const int entries = M.n_items * kWarps + 1;
    build_item_rows_kernel<int><<<(entries + threads - 1) / threads, threads, 0, stream>>>(M.rowPtr, M.rows, M.nnz, entries, M.item_row);
    launches++;
    const size_t part_bytes = sizeof(PartSlot) * (size_t)M.n_items * kWarps * 2;
    M.mean_len = M.rows > 0 ? (double)M.nnz / (double)M.rows : 0.0;
    M.len_cv = 0.0;
    if (M.rows > 0 && M.nnz > 0) {   // std/mean of the row lengths; the first packet slot is free scratch until it is armed below
        unsigned long long sq = 0;
        HPR_CUDA_CHECK(cudaMemsetAsync(M.head_part, 0, sizeof(PartSlot), stream));
        row_len_sq_kernel<<<std::min((M.rows + threads - 1) / threads, 1184), threads, 0, stream>>>(M.rowPtr, M.rows, M.head_part);
        launches++;
        HPR_CUDA_CHECK(cudaMemcpyAsync(&sq, M.head_part, sizeof(sq), cudaMemcpyDeviceToHost, stream));
        HPR_CUDA_CHECK(cudaStreamSynchronize(stream));
        const double var = (double)sq / (double)M.rows - M.mean_len * M.mean_len;
        M.len_cv = var > 0.0 ? std::sqrt(var) / M.mean_len : 0.0;
    }
    HPR_CUDA_CHECK(cudaMemsetAsync(M.head_part, 0xFF, part_bytes, stream));   // every packet "not published"
    HPR_CUDA_CHECK(cudaMemsetAsync(M.tail_part, 0xFF, part_bytes, stream));
    HPR_CUDA_CHECK(cudaMemsetAsync(M.ticket, 0, sizeof(unsigned), stream));
    M.tickets_issued = 0;
}

// Column bands of M for passes whose gathered vector (M.cols doubles) does not fit the L2: 8-byte gathers from DRAM run at
// 65 G/s against 278 G/s from the L2 (profiles/r1_gather_ceiling.json, 400 MB vs 8-40 MB vectors).  Each band is a CSR
// matrix over all rows with the entries of one column slice (global column indices, original order inside a row).
void Engine::build_bands(DevCsr &M) {
    long long band_cols = 0;
    if (const char *e = getenv("HPRLP_BAND_COLS")) band_cols = atoll(e);          // tests / tuning
    else if ((size_t)M.cols * sizeof(double) > ((size_t)64 << 20)) band_cols = 4 << 20;   // > 64 MB: 32 MB slices
    if (band_cols <= 0 || band_cols >= M.cols || M.nnz == 0) return;
    int nb = (int)((M.cols + band_cols - 1) / band_cols);
    if (nb > 64) { band_cols = (M.cols + 63) / 64; nb = (int)((M.cols + band_cols - 1) / band_cols); }
    const size_t stride = (size_t)M.rows + 1;
    int *brp = nullptr;
    HPR_CUDA_CHECK(cudaMalloc(&brp, sizeof(int) * stride * nb));
    HPR_CUDA_CHECK(cudaMemsetAsync(brp, 0, sizeof(int) * stride * nb, stream));
    std::vector<long long> bn(nb, 0);
    band_count(M.rows, M.rowPtr, M.col, (int)band_cols, nb, brp, bn.data(), stream);
    auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
    size_t total = 0;
    std::vector<size_t> o_col(nb), o_val(nb), o_item(nb), o_head(nb), o_tail(nb), o_tick(nb);
    std::vector<int> items(nb);
    for (int b = 0; b < nb; ++b) {
        items[b] = std::max(1, (int)((bn[b] + kChunk - 1) / kChunk));
        const size_t padded = (size_t)items[b] * kChunk, wit = (size_t)items[b] * kWarps;
        o_col[b] = total;  total += up(padded * sizeof(int));
        o_val[b] = total;  total += up(padded * sizeof(double));
        o_item[b] = total; total += up((wit + 1) * sizeof(int));
        o_head[b] = total; total += up(wit * 2 * sizeof(PartSlot));
        o_tail[b] = total; total += up(wit * 2 * sizeof(PartSlot));
        o_tick[b] = total; total += up(sizeof(unsigned));
    }
    const size_t o_carry = total; total += up((size_t)M.rows * sizeof(double));
    const size_t o_ptrs = total;  total += up(sizeof(void *) * 2 * nb);
    char *store = nullptr;
    HPR_CUDA_CHECK(cudaMalloc(&store, total));
    HPR_CUDA_CHECK(cudaMemsetAsync(store, 0, total, stream));   // padding entries (col 0, value 0); packets are armed by finish_matrix
    std::vector<void *> ptrs(2 * nb);
    M.bands.assign(nb, DevCsr());
    for (int b = 0; b < nb; ++b) {
        DevCsr &Bd = M.bands[b];
        Bd.rows = M.rows; Bd.cols = M.cols; Bd.nnz = bn[b]; Bd.n_items = items[b];
        Bd.rowPtr = brp + (size_t)b * stride;
        Bd.col = reinterpret_cast<int *>(store + o_col[b]);
        Bd.val = reinterpret_cast<double *>(store + o_val[b]);
        Bd.item_row = reinterpret_cast<int *>(store + o_item[b]);
        Bd.head_part = reinterpret_cast<PartSlot *>(store + o_head[b]);
        Bd.tail_part = reinterpret_cast<PartSlot *>(store + o_tail[b]);
        Bd.ticket = reinterpret_cast<unsigned *>(store + o_tick[b]);
        ptrs[b] = Bd.col; ptrs[nb + b] = Bd.val;
    }
    HPR_CUDA_CHECK(cudaMemcpyAsync(store + o_ptrs, ptrs.data(), sizeof(void *) * 2 * nb, cudaMemcpyHostToDevice, stream));
    band_fill(M.rows, M.rowPtr, M.col, M.val, (int)band_cols, nb, brp, reinterpret_cast<int *const *>(store + o_ptrs),
              reinterpret_cast<double *const *>(store + o_ptrs) + nb, stream);
    HPR_CUDA_CHECK(cudaStreamSynchronize(stream));   // ptrs (host vector) must outlive the copy
    for (int b = 0; b < nb; ++b) {
        finish_matrix(M.bands[b]);
        M.bands[b].G = pick_lanes(M.bands[b].mean_len, M.bands[b].len_cv, "HPRLP_LANES_BAND");
    }
    M.carry = reinterpret_cast<double *>(store + o_carry);
    M.band_store = store;
    M.band_rowptr_store = brp;
    launches += 2;
}

void Engine::alloc_common() {
    const size_t nv = dist() ? npad : (size_t)n;   // partitioned: exchange buffers hold nranks equal blocks
    if (dist()) px = peer_exchange_create(coll, device, npad, stream);   // collective: every rank, same point of the setup
    x = dalloc<double>(nv); x0 = dalloc<double>(nv); x_bar = dalloc<double>(nv);
    z_bar = dalloc<double>(nv); x_tmp = dalloc<double>(nv);
    // x_hat lives in peer-visible (cudaMalloc + IPC) memory when the NVLink exchange is on: the peers store their blocks into it
    x_hat = px ? px->xhat[rank] : dalloc<double>(nv);
    if (px) {   // the push pass of this rank starts at the x-block after its own (partial_ATy_pass)
        const size_t j = std::min<size_t>((size_t)n, xblock * (size_t)((rank + 1) % nranks));
        int p = 0;
        HPR_CUDA_CHECK(cudaMemcpyAsync(&p, AT.rowPtr + j, sizeof(int), cudaMemcpyDeviceToHost, stream));
        HPR_CUDA_CHECK(cudaStreamSynchronize(stream));
        AT.push_chunk0 = (int)(((long long)p / kChunk) % std::max(AT.n_items, 1));
    }
    wn = dalloc<double>(nv);
    y = dalloc<double>(m); y0 = dalloc<double>(m); y_bar = dalloc<double>(m); y_obj = dalloc<double>(m);
    y_tmp = dalloc<double>(m); wm = dalloc<double>(m); wm2 = dalloc<double>(m);
    row_norm = dalloc<double>(m); col_norm = dalloc<double>(n);
    d_params = dalloc<double>(4);
    d_k = dalloc<int>(2);
    partial_blocks = std::max(std::max(part_blocks(A), part_blocks(AT)), kVecBlocks);
    d_partials = dalloc<double>((size_t)partial_blocks * kMaxSlots);
    d_scal = dalloc<double>(16);
    h_scal = pinned_block_acquire();      // 16 residual / reduction scalars
    h_params = h_scal + 16;               // 4 sigma parameters (same recycled pinned block)
    int tex_limit = 0;   // linear textures hold at most this many (8-byte) elements; longer vectors are gathered with ld.global.nc
    HPR_CUDA_CHECK(cudaDeviceGetAttribute(&tex_limit, cudaDevAttrMaxTexture1DLinearWidth, device));
    if (const char *e = getenv("HPRLP_TEX_LIMIT")) tex_limit = atoi(e);   // tests: force the non-texture variants
    auto make_tex = [tex_limit](double *ptr, size_t count) -> cudaTextureObject_t {
        if (count > (size_t)tex_limit) return 0;
        cudaResourceDesc rd{}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = ptr;
        rd.res.linear.desc = cudaCreateChannelDesc<int2>(); rd.res.linear.sizeInBytes = count * sizeof(double);
        cudaTextureDesc td{}; td.readMode = cudaReadModeElementType;
        cudaTextureObject_t t = 0;
        HPR_CUDA_CHECK(cudaCreateTextureObject(&t, &rd, &td, nullptr));
        return t;
    };
    tex_y = make_tex(y, m); tex_xhat = make_tex(x_hat, n);
    tex_q = make_tex(wm2, m); tex_atq = make_tex(wn, n);   // power iteration: q and A^T q
    tex_ybar = make_tex(y_bar, m); tex_xbar = make_tex(x_bar, n); tex_xtmp = make_tex(x_tmp, n);   // check passes
    A.G = pick_lanes(A.mean_len, A.len_cv, "HPRLP_LANES_A");
    AT.G = pick_lanes(AT.mean_len, AT.len_cv, "HPRLP_LANES_AT");
}

// Host -> device copy of a large PAGEABLE array.  cudaMemcpyAsync from pageable memory is staged by the driver through one
// pinned buffer by one thread (~10-12 GB/s measured on the B200 boxes: 0.1 s for the 1.2 GB matrix of C3, most of the
// e2e setup time).  Here kUpThreads (8) host threads each stage their slice through two recycled 8 MB pinned buffers on their
// own stream (memcpy of chunk k+1 overlaps the DMA of chunk k).  The staging buffers live for the process (like the pinned
// scalar blocks); a second concurrent upload (partitioned mode: one host thread per GPU) falls back to the plain copy.
namespace {
constexpr int kUpThreads = 8;
constexpr size_t kUpChunk = (size_t)8 << 20;
std::mutex g_stage_mu;
char *g_stage[kUpThreads][2] = {};
}
void h2d_large(void *dst, const void *src, size_t bytes, cudaStream_t stream) {
    static const bool off = getenv("HPRLP_PLAIN_H2D") != nullptr;
    if (off || bytes < ((size_t)32 << 20) || !g_stage_mu.try_lock()) {
        HPR_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream));
        return;
    }
    std::lock_guard<std::mutex> lk(g_stage_mu, std::adopt_lock);
    for (int t = 0; t < kUpThreads; ++t)
        for (int b = 0; b < 2; ++b)
            if (!g_stage[t][b]) HPR_CUDA_CHECK(cudaHostAlloc(reinterpret_cast<void **>(&g_stage[t][b]), kUpChunk, cudaHostAllocPortable));
    cudaEvent_t ready;   // everything queued on the engine stream so far (the arena zero-fill) precedes the copies
    HPR_CUDA_CHECK(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
    HPR_CUDA_CHECK(cudaEventRecord(ready, stream));
    int dev = 0;
    cudaGetDevice(&dev);
    std::vector<std::thread> workers;
    std::vector<cudaError_t> errs(kUpThreads, cudaSuccess);
    for (int t = 0; t < kUpThreads; ++t) {
        workers.emplace_back([&, t]() {
            cudaSetDevice(dev);
            const size_t lo = (bytes * t / kUpThreads) & ~(size_t)255, hi = (t + 1 == kUpThreads) ? bytes : ((bytes * (t + 1) / kUpThreads) & ~(size_t)255);
            cudaStream_t st;
            cudaEvent_t ev[2];
            if ((errs[t] = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking)) != cudaSuccess) return;
            cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming);
            cudaStreamWaitEvent(st, ready, 0);
            int b = 0;
            bool used[2] = {false, false};
            for (size_t o = lo; o < hi; o += kUpChunk, b ^= 1) {
                const size_t len = std::min(kUpChunk, hi - o);
                if (used[b]) cudaEventSynchronize(ev[b]);
                std::memcpy(g_stage[t][b], static_cast<const char *>(src) + o, len);
                const cudaError_t e = cudaMemcpyAsync(static_cast<char *>(dst) + o, g_stage[t][b], len, cudaMemcpyHostToDevice, st);
                if (e != cudaSuccess) { errs[t] = e; break; }
                cudaEventRecord(ev[b], st);
                used[b] = true;
            }
            const cudaError_t e = cudaStreamSynchronize(st);
            if (errs[t] == cudaSuccess) errs[t] = e;
            cudaEventDestroy(ev[0]); cudaEventDestroy(ev[1]);
            cudaStreamDestroy(st);
        });
    }
    for (auto &w : workers) w.join();
    cudaEventDestroy(ready);
    for (cudaError_t e : errs) HPR_CUDA_CHECK(e);
}

// Device -> host copy of a large array into PAGEABLE (typically freshly malloc'ed) memory: the mirror image of h2d_large.
// The worker threads also take the first-touch page faults of the destination in parallel.
void d2h_large(void *dst, const void *src, size_t bytes, cudaStream_t stream) {
    static const bool off = getenv("HPRLP_PLAIN_H2D") != nullptr;
    if (off || bytes < ((size_t)32 << 20) || !g_stage_mu.try_lock()) {
        HPR_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, stream));
        HPR_CUDA_CHECK(cudaStreamSynchronize(stream));
        return;
    }
    std::lock_guard<std::mutex> lk(g_stage_mu, std::adopt_lock);
    for (int t = 0; t < kUpThreads; ++t)
        for (int b = 0; b < 2; ++b)
            if (!g_stage[t][b]) HPR_CUDA_CHECK(cudaHostAlloc(reinterpret_cast<void **>(&g_stage[t][b]), kUpChunk, cudaHostAllocPortable));
    cudaEvent_t ready;   // the kernels that produce src are queued on the engine stream
    HPR_CUDA_CHECK(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
    HPR_CUDA_CHECK(cudaEventRecord(ready, stream));
    int dev = 0;
    cudaGetDevice(&dev);
    std::vector<std::thread> workers;
    std::vector<cudaError_t> errs(kUpThreads, cudaSuccess);
    for (int t = 0; t < kUpThreads; ++t) {
        workers.emplace_back([&, t]() {
            cudaSetDevice(dev);
            const size_t lo = (bytes * t / kUpThreads) & ~(size_t)255, hi = (t + 1 == kUpThreads) ? bytes : ((bytes * (t + 1) / kUpThreads) & ~(size_t)255);
            cudaStream_t st;
            cudaEvent_t ev[2];
            if ((errs[t] = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking)) != cudaSuccess) return;
            cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming);
            cudaStreamWaitEvent(st, ready, 0);
            size_t pend_o[2] = {0, 0}, pend_len[2] = {0, 0};
            int b = 0;
            auto drain = [&](int bb) {   // chunk in staging buffer bb has arrived: hand it to the caller's array
                if (!pend_len[bb]) return;
                cudaEventSynchronize(ev[bb]);
                std::memcpy(static_cast<char *>(dst) + pend_o[bb], g_stage[t][bb], pend_len[bb]);
                pend_len[bb] = 0;
            };
            for (size_t o = lo; o < hi; o += kUpChunk, b ^= 1) {
                const size_t len = std::min(kUpChunk, hi - o);
                drain(b);
                const cudaError_t e = cudaMemcpyAsync(g_stage[t][b], static_cast<const char *>(src) + o, len, cudaMemcpyDeviceToHost, st);
                if (e != cudaSuccess) { errs[t] = e; break; }
                cudaEventRecord(ev[b], st);
                pend_o[b] = o; pend_len[b] = len;
            }
            drain(b); drain(b ^ 1);
            const cudaError_t e = cudaStreamSynchronize(st);
            if (errs[t] == cudaSuccess) errs[t] = e;
            cudaEventDestroy(ev[0]); cudaEventDestroy(ev[1]);
            cudaStreamDestroy(st);
        });
    }
    for (auto &w : workers) w.join();
    cudaEventDestroy(ready);
    for (cudaError_t e : errs) HPR_CUDA_CHECK(e);
}

// One zero-filled block from the engines' private pool (batched solver state), stream-ordered.
void *pool_alloc_zeroed(size_t bytes, int device, cudaStream_t st) {
    void *p = nullptr;
    static const bool no_pool = getenv("HPRLP_NO_POOL") != nullptr;
    if (no_pool) HPR_CUDA_CHECK(cudaMalloc(&p, bytes));
    else HPR_CUDA_CHECK(cudaMallocFromPoolAsync(&p, bytes, engine_pool(device), st));
    HPR_CUDA_CHECK(cudaMemsetAsync(p, 0, bytes, st));
    return p;
}
void *pool_alloc_raw(size_t bytes, int device, cudaStream_t st) {   // stream-ordered, not zeroed; release with cudaFreeAsync
    void *p = nullptr;
    HPR_CUDA_CHECK(cudaMallocFromPoolAsync(&p, std::max<size_t>(bytes, 16), engine_pool(device), st));
    return p;
}
void pool_free(void *p, cudaStream_t st) {
    if (!p) return;
    static const bool no_pool = getenv("HPRLP_NO_POOL") != nullptr;
    if (no_pool) { cudaFree(p); return; }
    cudaFreeAsync(p, st);
    cudaStreamSynchronize(st);
}

void Engine::upload(const LP_info_cpu *lp, int dev) {
    prepare(lp->m, lp->n, lp->A->numElements, dev);
    obj_constant = lp->obj_constant;
    HPR_CUDA_CHECK(cudaMemcpyAsync(A.rowPtr, lp->A->rowPtr, sizeof(int) * ((size_t)m + 1), cudaMemcpyHostToDevice, stream));
    h2d_large(A.col, lp->A->colIndex, sizeof(int) * (size_t)nnz, stream);
    h2d_large(A.val, lp->A->value, sizeof(double) * (size_t)nnz, stream);
    HPR_CUDA_CHECK(cudaMemcpyAsync(AL, lp->AL, sizeof(double) * m, cudaMemcpyHostToDevice, stream));
    HPR_CUDA_CHECK(cudaMemcpyAsync(AU, lp->AU, sizeof(double) * m, cudaMemcpyHostToDevice, stream));
    HPR_CUDA_CHECK(cudaMemcpyAsync(c, lp->c, sizeof(double) * n, cudaMemcpyHostToDevice, stream));
    HPR_CUDA_CHECK(cudaMemcpyAsync(l, lp->l, sizeof(double) * n, cudaMemcpyHostToDevice, stream));
    HPR_CUDA_CHECK(cudaMemcpyAsync(u, lp->u, sizeof(double) * n, cudaMemcpyHostToDevice, stream));
    static const bool host_transpose = getenv("HPRLP_HOST_TRANSPOSE") != nullptr;
    if (host_transpose) {
        std::vector<int> trp((size_t)n + 1), tci((size_t)nnz);
        std::vector<double> tv((size_t)nnz);
        csr_transpose_host(m, n, (int)nnz, lp->A->rowPtr, lp->A->colIndex, lp->A->value, trp.data(), tci.data(), tv.data());
        HPR_CUDA_CHECK(cudaMemcpyAsync(AT.rowPtr, trp.data(), sizeof(int) * ((size_t)n + 1), cudaMemcpyHostToDevice, stream));
        HPR_CUDA_CHECK(cudaMemcpyAsync(AT.col, tci.data(), sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice, stream));
        HPR_CUDA_CHECK(cudaMemcpyAsync(AT.val, tv.data(), sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice, stream));
        HPR_CUDA_CHECK(cudaStreamSynchronize(stream));
    }
    finish_setup(!host_transpose);
}

// Device, stream, the pooled arena and every problem buffer (A, A^T, AL, AU, c, l, u); the caller then fills A and
// the vectors (H2D copies in upload(), generator kernels in the synthetic partitioned path) and calls finish_setup().
void Engine::prepare(int m_, int n_, long long nnz_, int dev) {
    device = dev;
    HPR_CUDA_CHECK(cudaSetDevice(device));
    HPR_CUDA_CHECK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    m = m_; n = n_; nnz = nnz_;
    if (dist()) {   // x-block ownership: equal blocks of xblock columns (multiple of 64 entries = 512 B), the last one short
        xblock = (((size_t)n + nranks - 1) / nranks + 63) / 64 * 64;
        npad = xblock * nranks;
        xb0 = (int)std::min<size_t>((size_t)n, xblock * rank);
        xb1 = (int)std::min<size_t>((size_t)n, xblock * (rank + 1));
    } else {
        xblock = npad = (size_t)n; xb0 = 0; xb1 = n;
    }
    {
        // arena size: two padded CSR copies + item tables + 5 problem vectors + 9 n-vectors + 8 m-vectors + partials
        const size_t ctas = (size_t)((nnz + kChunk - 1) / kChunk) + 1, padded = ctas * kChunk, witems = ctas * kWarps;
        size_t need = 0;
        for (size_t rows : {(size_t)m, (size_t)n})
            need += arena_round((rows + 1) * 4) + arena_round(padded * 4) + arena_round(padded * 8) + arena_round((witems + 1) * 4) +
                    2 * arena_round(witems * 2 * sizeof(PartSlot)) + arena_round(8);
        need += 10 * arena_round((size_t)m * 8) + 13 * arena_round((dist() ? npad : (size_t)n) * 8);
        need += arena_round((size_t)(ctas + witems / kWarps + kVecBlocks + 64) * kMaxSlots * 8) + (1u << 16);
        Arena *ar = new Arena;
        ar->size = need;
        // Stream-ordered allocation from a PRIVATE memory pool per device (the process's default pool is left alone): the
        // arena of a finished solve stays cached up to the pool's release threshold, so repeated solve() calls pay neither
        // cudaMalloc nor cudaFree (both synchronous and ~0.1-0.5 s for multi-GB buffers).  HPRLP_POOL_RETAIN_MB bounds what
        // stays cached (default 8192 MB; 0 = return everything as soon as an engine is destroyed);
        // hprlp_b200_release_cached_memory() returns it all on demand.  HPRLP_NO_POOL=1 restores cudaMalloc/cudaFree.
        static const bool no_pool = getenv("HPRLP_NO_POOL") != nullptr;
        pooled_ = !no_pool;
        if (pooled_) {
            HPR_CUDA_CHECK(cudaMallocFromPoolAsync(reinterpret_cast<void **>(&ar->base), need, engine_pool(device), stream));
        } else {
            HPR_CUDA_CHECK(cudaMalloc(&ar->base, need));
        }
        HPR_CUDA_CHECK(cudaMemsetAsync(ar->base, 0, need, stream));
        arena_ = ar;
        g_arena = ar;
    }
    alloc_matrix(A, m, n, nnz);
    alloc_matrix(AT, n, m, nnz);
    AL = dalloc<double>(m); AU = dalloc<double>(m); c = dalloc<double>(n); l = dalloc<double>(n); u = dalloc<double>(n);
}

// A^T on the device in the reference's entry order (stable sort by column, transpose.cu), item tables, work vectors.
void Engine::finish_setup(bool build_transpose) {
    if (build_transpose) device_transpose_csr(m, n, (int)nnz, A.rowPtr, A.col, A.val, AT.rowPtr, AT.col, AT.val, stream);
    finish_matrix(A);
    finish_matrix(AT);
    alloc_common();
    zo_buf = dalloc<double>(dist() ? npad : (size_t)n);
    g_arena = nullptr;
    HPR_CUDA_CHECK(cudaStreamSynchronize(stream));
}

// plain SpMV helpers for setup-time use (synthetic instance generation): out = A g / out = A^T g
void Engine::spmv_A(const double *g, double *out) {
    SpmvOp<false> o; o.g = g; o.tex = 0; o.out = out; o.q = nullptr; o.partials = nullptr;
    launch_stream(A, o, stream);
}
void Engine::spmv_AT(const double *g, double *out) {
    SpmvOp<false> o; o.g = g; o.tex = 0; o.out = out; o.q = nullptr; o.partials = nullptr;
    launch_stream(AT, o, stream);
}

Engine::~Engine() {
    static const bool timing = getenv("HPRLP_TIMING") != nullptr;
    double t[6] = {0, 0, 0, 0, 0, 0};
    t[0] = now_seconds();
    if (stream) cudaStreamSynchronize(stream);
    for (auto &kv : graphs_) cudaGraphExecDestroy(kv.second);
    graphs_.clear();
    t[1] = now_seconds();
    for (cudaTextureObject_t tx : {tex_y, tex_xhat, tex_q, tex_atq, tex_ybar, tex_xbar, tex_xtmp})
        if (tx) cudaDestroyTextureObject(tx);
    if (px) { peer_exchange_destroy(px, coll, stream); px = nullptr; }   // collective: agrees that no peer still uses the buffers
    t[2] = now_seconds();
    for (DevCsr *M : {&A, &AT}) {
        if (M->band_store) cudaFree(M->band_store);
        if (M->band_rowptr_store) cudaFree(M->band_rowptr_store);
        M->bands.clear();
    }
    if (arena_) {   // every device buffer of this engine lives in the arena (the column bands above excepted)
        Arena *ar = static_cast<Arena *>(arena_);
        if (pooled_ && stream) { cudaFreeAsync(ar->base, stream); cudaStreamSynchronize(stream); }
        else cudaFree(ar->base);
        delete ar;
    }
    t[3] = now_seconds();
    pinned_block_release(h_scal);         // h_params lives in the same block
    t[4] = now_seconds();
    if (stream) cudaStreamDestroy(stream);
    t[5] = now_seconds();
    if (timing)
        fprintf(stderr, "[hprlp timing] teardown: graphs %.4f, textures %.4f, arena %.4f, pinned %.4f, stream %.4f s\n", t[1] - t[0],
                t[2] - t[1], t[3] - t[2], t[4] - t[3], t[5] - t[4]);
}

void Engine::set_partition(Collective *c, int m_global_, int row0_) {
    coll = c;
    nranks = c ? c->nranks : 1;
    rank = c ? c->rank : 0;
    m_global = m_global_;
    row0 = row0_;
}

void Engine::allreduce(double *buf, size_t count, bool max_op) {
    if (!coll) return;
    coll->all_reduce(buf, count, max_op, stream);
}

void Engine::fetch_scalars(int count) {
    HPR_CUDA_CHECK(cudaMemcpyAsync(h_scal, d_scal, sizeof(double) * count, cudaMemcpyDeviceToHost, stream));
    HPR_CUDA_CHECK(cudaStreamSynchronize(stream));
}

// ------------------------------------------------------------------------------------------------
// scaling (reference src/scaling.cu:88-216)
// ------------------------------------------------------------------------------------------------
void Engine::scale(const HPRLP_parameters *p) {
    double *t1 = wm, *t2 = wn;
    const int gm = vec_grid(m), gn = vec_grid(n);
    fill_kernel<<<gm, kVecThreads, 0, stream>>>(row_norm, m, 1.0);
    fill_kernel<<<gn, kVecThreads, 0, stream>>>(col_norm, n, 1.0);
    launches += 2;

    auto norms_bc = [&](double *nb, double *nc) {
        norm_bc_kernel<<<kVecBlocks, kVecThreads, 0, stream>>>(AL, AU, m, c, n, d_partials);
        final_reduce_kernel<<<1, 1024, 0, stream>>>(d_partials, kVecBlocks, 2, d_scal);
        launches += 2;
        allreduce(d_scal, 1);   // |b|^2 is a sum over (partitioned) rows; |c|^2 is replicated
        fetch_scalars(2);
        *nb = sqrt(h_scal[0]);
        *nc = sqrt(h_scal[1]);
    };
    double nb, nc;
    norms_bc(&nb, &nc);
    norm_b_org = 1.0 + nb;
    norm_c_org = 1.0 + nc;

    const CsrView<int> vA = view_of(A), vAT = view_of(AT);
    if (p->use_CR_scaling) {
        // 20 alternating log-mean sweeps from zero (src/scaling.cu:40-64)
        fill_kernel<<<gm, kVecThreads, 0, stream>>>(t1, m, 0.0);
        fill_kernel<<<gn, kVecThreads, 0, stream>>>(t2, n, 0.0);
        launches += 2;
        for (int it = 0; it < 20; ++it) {
            CurtisReidOp<false> oa; oa.other = t2; oa.out = t1; oa.cnt_out = nullptr;
            launch_stream(A, oa, stream);
            if (!dist()) {
                CurtisReidOp<false> ob; ob.other = t1; ob.out = t2; ob.cnt_out = nullptr;
                launch_stream(AT, ob, stream);
            } else {   // column means over ALL rows: local sums and counts, cross-GPU sum, divide
                CurtisReidOp<true> ob; ob.other = t1; ob.out = t2; ob.cnt_out = x_tmp;
                launch_stream(AT, ob, stream);
                allreduce(t2, n); allreduce(x_tmp, n);
                mean_kernel<<<gn, kVecThreads, 0, stream>>>(t2, x_tmp, n);
            }
            launches += 2;
        }
        exp_clamp_kernel<<<gm, kVecThreads, 0, stream>>>(t1, m);
        exp_clamp_kernel<<<gn, kVecThreads, 0, stream>>>(t2, n);
        scale_values_kernel<false, true, int><<<A.n_items, kThreads, 0, stream>>>(vA, A.val, t1, t2);
        scale_values_kernel<false, false, int><<<AT.n_items, kThreads, 0, stream>>>(vAT, AT.val, t2, t1);
        scale_row_vectors_kernel<true><<<gm, kVecThreads, 0, stream>>>(t1, row_norm, AL, AU, m);
        scale_col_vectors_kernel<true><<<gn, kVecThreads, 0, stream>>>(t2, col_norm, c, l, u, n);
        launches += 6;
    }
    const int ruiz_rounds = p->use_Ruiz_scaling ? 10 : 0;
    const int rounds = ruiz_rounds + (p->use_Pock_Chambolle_scaling ? 1 : 0);
    for (int it = 0; it < rounds; ++it) {
        // both statistics are taken from the matrix before either factor is applied (:127-144)
        const bool is_max = it < ruiz_rounds;
        if (is_max) { RowNormOp<true, false> oa; oa.out = t1; launch_stream(A, oa, stream); }
        else        { RowNormOp<false, false> oa; oa.out = t1; launch_stream(A, oa, stream); }
        if (!dist()) {
            if (is_max) { RowNormOp<true, false> ob; ob.out = t2; launch_stream(AT, ob, stream); }
            else        { RowNormOp<false, false> ob; ob.out = t2; launch_stream(AT, ob, stream); }
        } else {       // column statistic over ALL rows: raw local max / sum, cross-GPU reduce, then sqrt + clamp
            if (is_max) { RowNormOp<true, true> ob; ob.out = t2; launch_stream(AT, ob, stream); }
            else        { RowNormOp<false, true> ob; ob.out = t2; launch_stream(AT, ob, stream); }
            allreduce(t2, n, is_max);
            sqrt_clamp_kernel<<<gn, kVecThreads, 0, stream>>>(t2, n);
        }
        scale_values_kernel<true, true, int><<<A.n_items, kThreads, 0, stream>>>(vA, A.val, t1, t2);
        scale_values_kernel<true, false, int><<<AT.n_items, kThreads, 0, stream>>>(vAT, AT.val, t2, t1);
        scale_row_vectors_kernel<false><<<gm, kVecThreads, 0, stream>>>(t1, row_norm, AL, AU, m);
        scale_col_vectors_kernel<false><<<gn, kVecThreads, 0, stream>>>(t2, col_norm, c, l, u, n);
        launches += 6;
    }
    if (p->use_bc_scaling) {
        norms_bc(&nb, &nc);
        b_scale = 1.0 + nb;
        c_scale = 1.0 + nc;
        const double bs = 1.0 / b_scale, cs = 1.0 / c_scale;
        scal_kernel<<<gm, kVecThreads, 0, stream>>>(AU, m, bs);
        scal_kernel<<<gm, kVecThreads, 0, stream>>>(AL, m, bs);
        scal_kernel<<<gn, kVecThreads, 0, stream>>>(l, n, bs);
        scal_kernel<<<gn, kVecThreads, 0, stream>>>(u, n, bs);
        scal_kernel<<<gn, kVecThreads, 0, stream>>>(c, n, cs);
        launches += 5;
    } else {
        b_scale = 1.0;
        c_scale = 1.0;
    }
    norms_bc(&nb, &nc);
    norm_b = nb;
    norm_c = nc;
    HPR_CUDA_CHECK(cudaGetLastError());
    // after scaling (the bands copy the final values); no-ops unless the gathered vector exceeds the L2 budget
    build_bands(A);    // passes over A gather n-vectors (x_hat, x_bar, A^T q)
    build_bands(AT);   // passes over A^T gather m-vectors (y, y_bar, q)
}

// ------------------------------------------------------------------------------------------------
// power iteration (reference src/power_iteration.cu:20-119)
// ------------------------------------------------------------------------------------------------
// First-solve warm-up (api.cu, hprlp_b200_warmup): creating the CUDA context (~0.3-2 s with the module load) and loading
// cuRAND's device code (~0.6-1.2 s for the first generator on B200) are paid once per process.  Both can run on a second
// host thread while this one parses the MPS file or runs the PSLP presolve.  Errors are ignored here -- the real calls
// report them.
void warm_device(int device) {
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return; }
    cudaFree(nullptr);   // context + module load
    double *buf = nullptr;
    if (cudaMalloc(&buf, 2 * sizeof(double)) != cudaSuccess) { cudaGetLastError(); return; }
    try { draw_power_start(device, buf, 2, nullptr); } catch (...) {}
    cudaFree(buf);
    try { engine_pool(device); } catch (...) {}
    cudaGetLastError();
}

void Engine::power_start_vector(double *d_z) {
    // cuRAND XORWOW (CURAND_RNG_PSEUDO_DEFAULT), seed 1, N(0,1), then + 1e-8.  For odd m the reference's
    // unchecked curandGenerateNormalDouble fails with LENGTH_NOT_MULTIPLE and leaves z = 0 (its Ax
    // buffer after the warm-up SpMV with x_bar = 0, src/preprocess.cu:153-156) => z = 1e-8 * ones.
    const int mg = dist() ? m_global : m;
    double *gen_buf = d_z;
    if (dist()) {   // every GPU draws the global vector (same seed) and keeps its own row block
        HPR_CUDA_CHECK(cudaMalloc(&gen_buf, sizeof(double) * (size_t)mg));
    }
    HPR_CUDA_CHECK(cudaMemsetAsync(gen_buf, 0, sizeof(double) * mg, stream));
    if ((mg % 2) == 0) draw_power_start(device, gen_buf, (size_t)mg, stream);
    if (dist()) {
        HPR_CUDA_CHECK(cudaMemcpyAsync(d_z, gen_buf + row0, sizeof(double) * m, cudaMemcpyDeviceToDevice, stream));
        HPR_CUDA_CHECK(cudaStreamSynchronize(stream));
        cudaFree(gen_buf);
    }
    add_scalar_kernel<<<vec_grid(m), kVecThreads, 0, stream>>>(d_z, m, 1e-8);
    launches++;
}

double Engine::power_iteration(int max_iter, double tol, const double *host_z0, int *iters_out) {
    double *z = wm, *q = wm2, *atq = wn;
    static const bool timing = getenv("HPRLP_TIMING") != nullptr;
    double t_start = 0.0;
    if (timing) { cudaStreamSynchronize(stream); t_start = now_seconds(); }
    if (host_z0) {
        HPR_CUDA_CHECK(cudaMemcpyAsync(z, host_z0, sizeof(double) * m, cudaMemcpyHostToDevice, stream));
    } else {
        power_start_vector(z);
    }
    cudaEvent_t pe0 = nullptr, pe1 = nullptr;
    double t_sync = 0.0, t_sync_max = 0.0;
    if (timing) {
        cudaStreamSynchronize(stream);
        fprintf(stderr, "[hprlp timing] power start vector %.4f s\n", now_seconds() - t_start);
        t_start = now_seconds();
        cudaEventCreate(&pe0); cudaEventCreate(&pe1);
        cudaEventRecord(pe0, stream);
    }
    // d_scal[0] = <z,z>, d_scal[1] = <q,z>, d_scal[2] = |z - lambda q|^2
    sumsq_kernel<<<kVecBlocks, kVecThreads, 0, stream>>>(z, m, d_partials);
    final_reduce_kernel<<<1, 1024, 0, stream>>>(d_partials, kVecBlocks, 1, d_scal);
    launches += 2;
    allreduce(d_scal, 1);
    double lambda = 1.0;
    auto one_iteration = [&]() {
        power_normalize_kernel<<<vec_grid(m), kVecThreads, 0, stream>>>(z, q, d_scal, m);
        if (dist() && push_mode()) {   // A^T q = sum over row blocks: rows pushed to their owners, summed there, sent to everyone's x_hat
            partial_ATy_pass(q, tex_q);
            exchange_x(false, 1);
            launch_spmv_hot<true>(A, x_hat, tex_xhat, z, q, d_partials, stream);
            launches += 3;
        } else {
            launch_spmv_hot<false>(AT, q, tex_q, atq, nullptr, nullptr, stream);
            allreduce(atq, n);   // A^T q = sum over row blocks
            launch_spmv_hot<true>(A, atq, tex_atq, z, q, d_partials, stream);
        }
        final_reduce_kernel<<<1, 1024, 0, stream>>>(d_partials, part_blocks(A), 2, d_scal);
        allreduce(d_scal, 2);
        launches += 4;
    };
    auto error_estimate = [&]() {   // every 10th iteration, like the reference
        power_error_kernel<<<kVecBlocks, kVecThreads, 0, stream>>>(z, q, d_scal + 1, m, d_partials);
        final_reduce_kernel<<<1, 1024, 0, stream>>>(d_partials, kVecBlocks, 1, d_scal + 2);
        launches += 2;
        allreduce(d_scal + 2, 1);
    };
    // Small matrices are launch-bound here (configs[3]: 1640 iterations of four 5-10 us kernels): ten iterations and their
    // error estimate replay as one graph once the first 30 have run directly (an LP that converges at once never pays the
    // instantiation).  Same kernels, same arguments, same order: the result does not depend on the path.
    static const bool no_graph = getenv("HPRLP_NO_GRAPH") != nullptr;
    const bool graph_ok = !no_graph && !dist() && nnz <= 4000000 && A.bands.empty() && AT.bands.empty();
    constexpr int kPowerGraphKey = -10;   // graphs_ keys > 0 are loop graphs (run_normal)
    int it = 0;
    bool converged = false;
    while (it < max_iter) {
        const int blk = std::min(10, max_iter - it);
        if (blk == 10 && graph_ok && it >= 30) {
            auto g = graphs_.find(kPowerGraphKey);
            if (g == graphs_.end()) {
                cudaGraph_t gr = nullptr;
                cudaGraphExec_t ge = nullptr;
                HPR_CUDA_CHECK(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
                const long long before = launches;
                for (int i = 0; i < 10; ++i) one_iteration();
                error_estimate();
                launches = before;
                HPR_CUDA_CHECK(cudaStreamEndCapture(stream, &gr));
                HPR_CUDA_CHECK(cudaGraphInstantiate(&ge, gr, nullptr, nullptr, 0));
                cudaGraphDestroy(gr);
                g = graphs_.emplace(kPowerGraphKey, ge).first;
            }
            for (DevCsr *M : {&A, &AT}) {   // ten launches per matrix inside the graph: keep the ticket bookkeeping exact
                if (M->tickets_issued + 10ull * M->n_items > ticket_wrap()) {
                    HPR_CUDA_CHECK(cudaMemsetAsync(M->ticket, 0, sizeof(unsigned), stream));
                    M->tickets_issued = 0;
                }
                M->tickets_issued += 10ull * M->n_items;
            }
            HPR_CUDA_CHECK(cudaGraphLaunch(g->second, stream));
            launches += 42;
        } else {
            for (int i = 0; i < blk; ++i) one_iteration();
            if (blk == 10) error_estimate();
        }
        it += blk;
        if (blk == 10) {
            const double ts0 = timing ? now_seconds() : 0.0;
            fetch_scalars(3);
            if (timing) { const double d = now_seconds() - ts0; t_sync += d; t_sync_max = std::max(t_sync_max, d); }
            lambda = h_scal[1];
            if (sqrt(h_scal[2]) < tol) { converged = true; break; }
        }
    }
    if (!converged) {
        printf("Power iteration did not converge within the specified tolerance.\nMax iter: %d, Error: %.2e\n", max_iter,
               sqrt(h_scal[2]));
    }
    if (iters_out) *iters_out = it;
    HPR_CUDA_CHECK(cudaGetLastError());
    if (timing) {
        float dev_ms = 0.f;
        cudaEventRecord(pe1, stream); cudaEventSynchronize(pe1); cudaEventElapsedTime(&dev_ms, pe0, pe1);
        cudaEventDestroy(pe0); cudaEventDestroy(pe1);
        fprintf(stderr, "[hprlp timing] power loop %d iterations: wall %.4f s, device %.4f s, in syncs %.4f s (max %.4f s)\n", it,
                now_seconds() - t_start, dev_ms * 1e-3, t_sync, t_sync_max);
    }
    return lambda;
}

// ------------------------------------------------------------------------------------------------
// iteration building blocks
// ------------------------------------------------------------------------------------------------
void Engine::init_iterates() {
    for (double *v : {x, x0, x_hat, x_bar, z_bar, x_tmp}) HPR_CUDA_CHECK(cudaMemsetAsync(v, 0, sizeof(double) * (dist() ? npad : (size_t)n), stream));
    for (double *v : {y, y0, y_bar, y_obj, y_tmp}) HPR_CUDA_CHECK(cudaMemsetAsync(v, 0, sizeof(double) * m, stream));
    reset_halpern_counter();
    upload_params();
}

void Engine::upload_params() {   // reference reset_/upload_halpern_*_params, src/main_iterate.cu:17-66
    set_params_kernel<<<1, 1, 0, stream>>>(d_params, sigma, lambda_max);
    launches++;
}

void Engine::reset_halpern_counter() { HPR_CUDA_CHECK(cudaMemsetAsync(d_k, 0, 2 * sizeof(int), stream)); }

// The two fused passes of one HPR iteration.  Gathers go through the TEX pipe when the gathered vector could be bound as
// a linear texture (alloc_common), through ld.global.nc otherwise.
void Engine::launch_x_phase(bool check) {
    auto go = [&](auto ox) {
        ox.y = y; ox.tex = tex_y; ox.x = x; ox.x_hat = x_hat; ox.c = c; ox.l = l; ox.u = u; ox.x0 = x0;
        ox.x_bar = check ? x_bar : nullptr; ox.z_bar = check ? z_bar : nullptr; ox.x_tmp = check ? x_tmp : nullptr;
        ox.params = d_params; ox.kx = d_k; ox.ky = d_k + 1;
        launch_stream_hot(AT, ox, stream);
    };
    if (tex_y) { if (check) go(XPhaseOp<true, true>()); else go(XPhaseOp<false, true>()); }
    else       { if (check) go(XPhaseOp<true, false>()); else go(XPhaseOp<false, false>()); }
}
void Engine::launch_y_phase(bool check) {
    auto go = [&](auto oy) {
        oy.x_hat = x_hat; oy.tex = tex_xhat; oy.y = y; oy.AL = AL; oy.AU = AU; oy.y0 = y0;
        oy.y_bar = check ? y_bar : nullptr; oy.y_obj = check ? y_obj : nullptr; oy.y_tmp = check ? y_tmp : nullptr;
        oy.params = d_params; oy.ky = d_k + 1; oy.kx = d_k;
        launch_stream_hot(A, oy, stream);
    };
    if (tex_xhat) { if (check) go(YPhaseOp<true, true>()); else go(YPhaseOp<false, true>()); }
    else          { if (check) go(YPhaseOp<true, false>()); else go(YPhaseOp<false, false>()); }
}
// wn holds this rank's partial A_p^T y_p.  On return x (owned block) is updated and x_hat is complete on every rank.
void Engine::exchange_x(bool check, int mode, const double *src) {
    const int gx = vec_grid(xb1 - xb0);
    if (push_mode()) {   // the partials are already in (or on their way to) the owners' receive slots: partial_ATy_pass()
        PeerPtrs pp;
        for (int q = 0; q < kMaxPeers; ++q) { pp.xhat[q] = px->xhat[q]; pp.flags[q] = px->flags[q]; }
        const unsigned long long e = ++px->epoch;
        exchange_signal_kernel<<<1, 32, 0, stream>>>(pp, nranks, rank, 0, e);   // "my pass has delivered everywhere"
        const int gp = vec_grid((xb1 - xb0 + 1) / 2);   // one element pair per thread and pass
        auto launch = [&](auto kernel, bool chk) {
            kernel<<<gp, kVecThreads, 0, stream>>>(pp, mode == 3 ? src + xb0 : px->w[rank], xblock, nranks, rank, e, px->done, x, c, l, u, x0, chk ? x_bar : nullptr,
                                                   chk ? z_bar : nullptr, chk ? x_tmp : nullptr, d_params, d_k, d_k + 1, xb0, xb1);
        };
        if (mode == 1) launch(fused_exchange_x_kernel<false, 0, 1>, false);        // sum of the pushed partials to every rank's x_hat
        else if (mode == 2) launch(fused_exchange_x_kernel<false, 0, 2>, false);   // ... to my own x_hat block
        else if (mode == 3) launch(fused_exchange_x_kernel<false, 0, 3>, false);   // src block to every rank's x_hat
        else if (check) launch(fused_exchange_x_kernel<true, 0>, true);   // check iterations are rare: one generic instantiation
        else switch (nranks) {
            case 2: launch(fused_exchange_x_kernel<false, 2>, false); break;
            case 3: launch(fused_exchange_x_kernel<false, 3>, false); break;
            case 4: launch(fused_exchange_x_kernel<false, 4>, false); break;
            case 5: launch(fused_exchange_x_kernel<false, 5>, false); break;
            case 6: launch(fused_exchange_x_kernel<false, 6>, false); break;
            case 7: launch(fused_exchange_x_kernel<false, 7>, false); break;
            case 8: launch(fused_exchange_x_kernel<false, 8>, false); break;
            default: launch(fused_exchange_x_kernel<false, 0>, false); break;
        }
        if (mode != 2) exchange_wait_kernel<<<1, 32, 0, stream>>>(px->flags[rank], nranks, 1, e);
        launches += 3;
        return;
    }
    coll->reduce_scatter_inplace(wn, xblock, stream);
    if (check) x_update_kernel<true><<<gx, kVecThreads, 0, stream>>>(wn, x, x_hat, c, l, u, x0, x_bar, z_bar, x_tmp, d_params, d_k, d_k + 1, xb0, xb1);
    else x_update_kernel<false><<<gx, kVecThreads, 0, stream>>>(wn, x, x_hat, c, l, u, x0, nullptr, nullptr, nullptr, d_params, d_k, d_k + 1, xb0, xb1);
    coll->all_gather_inplace(x_hat, xblock, stream);
    launches += 1;
}

// The x-side pass of a row-partitioned iteration: w_p = A_p^T y_p.  NCCL transport: into wn (reduce-scattered afterwards).
// Peer-memory transport: every row result goes straight to the receive slot of the column's owner (SpmvPushOp), each rank
// starting at the x-block after its own so that the owners are targeted by one peer at a time.
void Engine::partial_ATy_pass(const double *g, cudaTextureObject_t tex) {
    if (!g) { g = y; tex = tex_y; }
    if (!push_mode()) { launch_spmv_hot<false>(AT, g, tex, wn, nullptr, nullptr, stream); return; }
    SpmvPushOp o;
    o.g = g; o.tex = tex; o.xblock = (int)xblock;
    for (int q = 0; q < kMaxPeers; ++q)
        o.slot[q] = q < nranks ? px->w[q] + ((long long)rank - q) * (long long)xblock : nullptr;
    CsrView<int> v = view_of(AT);
    v.chunk_offset = AT.push_chunk0;
    dispatch_hot(v, AT.G, o, stream);
}

void Engine::launch_iteration(bool check) {
    if (dist()) {
        // Row-partitioned iteration: partial w_p = A_p^T y_p over the local rows, reduce-scatter so that this
        // GPU holds (A^T y) on ITS x-block, x-update on that block only, all-gather of the x_hat blocks, then the fused
        // y-phase on the local rows of A with the full x_hat.
        partial_ATy_pass();
        exchange_x(check);
        launch_y_phase(check);
        launches += 2;
        return;
    }
    launch_x_phase(check);
    launch_y_phase(check);
    launches += 2;
}

// `count` normal iterations.  Kernel arguments are immutable (sigma and the Halpern counter live in
// device memory), so a captured graph of L iterations is replayable (the reference captures one
// iteration, src/HPRLP.cu:99-114); graphs are cached per length.
void Engine::run_normal(int count) {
    static const bool no_graph = getenv("HPRLP_NO_GRAPH") != nullptr;
    while (count > 0) {
        // Graph lengths are restricted to {8, 64} (at most two instantiations per solve: a 64-iteration graph has 256
        // kernel nodes); the remainder (< 8 iterations) is launched directly.
        int len = count >= 64 ? 64 : (count >= 8 ? 8 : count);
        // Graphs only pay when the loop is launch-bound (small matrices) and long enough to amortise the instantiation
        // (~30 ms for 256 nodes): matrices with >= 2e6 nonzeros keep the host ahead of the GPU with direct launches.
        const bool want_graph = nnz < 2000000 && loop.iter >= 300;
        if (no_graph || len < 8 || dist() || !want_graph) {
            for (int i = 0; i < len; ++i) launch_iteration(false);
        } else {
            auto it = graphs_.find(len);
            if (it == graphs_.end()) {
                cudaGraph_t g = nullptr;
                cudaGraphExec_t ge = nullptr;
                HPR_CUDA_CHECK(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
                const long long before = launches;
                for (int i = 0; i < len; ++i) launch_iteration(false);
                launches = before;
                HPR_CUDA_CHECK(cudaStreamEndCapture(stream, &g));
                HPR_CUDA_CHECK(cudaGraphInstantiate(&ge, g, nullptr, nullptr, 0));
                cudaGraphDestroy(g);
                it = graphs_.emplace(len, ge).first;
            }
            for (DevCsr *M : {&A, &AT}) {   // one launch per matrix and iteration inside the graph: keep the ticket bookkeeping exact
                if (M->tickets_issued + (unsigned long long)len * M->n_items > ticket_wrap()) {
                    HPR_CUDA_CHECK(cudaMemsetAsync(M->ticket, 0, sizeof(unsigned), stream));
                    M->tickets_issued = 0;
                }
                M->tickets_issued += (unsigned long long)len * M->n_items;
            }
            HPR_CUDA_CHECK(cudaGraphLaunch(it->second, stream));
            launches += 2LL * len;
        }
        count -= len;
    }
}

// reference compute_residuals, src/main_iterate.cu:229-309 -- two fused passes, one D2H of 9 scalars.
namespace { struct ResidTiming { double ms = 0.0, wait_ms = 0.0; int n = 0; } g_rt; }
void Engine::compute_residuals(int iter, bool compute_gap, Residuals *res, RestartState *rs) {
    static const bool timing = getenv("HPRLP_TIMING") != nullptr;
    cudaEvent_t te0 = nullptr, te1 = nullptr;
    double host0 = 0.0;
    if (timing) { cudaEventCreate(&te0); cudaEventCreate(&te1); cudaEventRecord(te0, stream); host0 = now_seconds(); }
    auto fill_dual = [&](auto &o) {
        o.y_bar = y_bar; o.c = c; o.z_bar = z_bar; o.x_bar = x_bar; o.x_tmp = x_tmp; o.col_norm = col_norm;
        o.l = l; o.u = u; o.tex = tex_ybar; o.partials = d_partials;
    };
    const bool push = dist() && push_mode();
    const double *xbar_g = x_bar, *xtmp_g = x_tmp;           // what the primal / gap passes gather from
    cudaTextureObject_t tex_xbar_g = tex_xbar, tex_xtmp_g = tex_xtmp;
    if (dist()) {
        const int gx = vec_grid(xb1 - xb0);
        const double *w = wn;
        if (push) {   // (A^T y_bar) on the owned block: rows pushed to their owners, summed into my x_hat block (free scratch here)
            partial_ATy_pass(y_bar, tex_ybar);
            exchange_x(false, 2);
            w = x_hat;
        } else {
            SpmvOp<false> ow; ow.g = y_bar; ow.tex = 0; ow.out = wn; ow.q = nullptr; ow.partials = nullptr;
            launch_stream(AT, ow, stream);
            coll->reduce_scatter_inplace(wn, xblock, stream);
        }
        if (iter == 0) residual_dual_kernel<false, true><<<gx, kVecThreads, 0, stream>>>(w, c, z_bar, x_bar, x_tmp, col_norm, l, u, xb0, xb1, d_partials);
        else if (compute_gap) residual_dual_kernel<true, false><<<gx, kVecThreads, 0, stream>>>(w, c, z_bar, x_bar, x_tmp, col_norm, l, u, xb0, xb1, d_partials);
        else residual_dual_kernel<false, false><<<gx, kVecThreads, 0, stream>>>(w, c, z_bar, x_bar, x_tmp, col_norm, l, u, xb0, xb1, d_partials);
        final_reduce_kernel<<<1, 1024, 0, stream>>>(d_partials, gx, 5, d_scal);   // sums over the owned x-block: all-reduced below
        if (push) {   // the primal pass gathers the full x_bar: every rank stores its block into everyone's x_hat
            exchange_x(false, 3, x_bar);
            xbar_g = x_hat; tex_xbar_g = tex_xhat;
        } else {
            coll->all_gather_inplace(x_bar, xblock, stream);
            if (compute_gap) coll->all_gather_inplace(x_tmp, xblock, stream);
        }
    } else {
    if (iter == 0) { ResidualDualOp<false, true> o; fill_dual(o); launch_stream(AT, o, stream); }
    else if (compute_gap) { ResidualDualOp<true, false> o; fill_dual(o); launch_stream(AT, o, stream); }
    else { ResidualDualOp<false, false> o; fill_dual(o); launch_stream(AT, o, stream); }
    final_reduce_kernel<<<1, 1024, 0, stream>>>(d_partials, part_blocks(AT), 5, d_scal);
    }
    auto fill_primal = [&](auto &o) {
        o.x_bar = xbar_g; o.x_tmp = x_tmp; o.AL = AL; o.AU = AU; o.row_norm = row_norm; o.y_obj = y_obj;
        o.y_bar = y_bar; o.y_tmp = y_tmp; o.tex = tex_xbar_g; o.partials = d_partials;
    };
    {
        ResidualPrimalOp<false> o; fill_primal(o); launch_stream(A, o, stream);
        final_reduce_kernel<<<1, 1024, 0, stream>>>(d_partials, part_blocks(A), 4, d_scal + 5);
        launches += 4;
    }
    if (compute_gap) {
        // restart gap terms <A dx, dy>, |dy|^2 (slots 7, 8) as a second single-product pass.  The two-product variant
        // (ResidualPrimalOp<true>) doubles the shared-memory staging; at 6 CTAs/SM that leaves ~28 KB of L1, i.e. hardly any
        // in-flight gather misses: measured 4.5 ms on C3 against 0.6 ms per single-product pass (same row sums bit for bit).
        if (push) { exchange_x(false, 3, x_tmp); xtmp_g = x_hat; tex_xtmp_g = tex_xhat; }   // all-gather of x_bar - x_hat (after the primal pass has read x_hat)
        WeightedNormOp o; o.dx = xtmp_g; o.dy = y_tmp; o.tex = tex_xtmp_g; o.partials = d_partials;
        launch_stream(A, o, stream);
        final_reduce_kernel<<<1, 1024, 0, stream>>>(d_partials, part_blocks(A), 2, d_scal + 7);
        launches += 2;
    }
    if (dist() && !compute_gap) HPR_CUDA_CHECK(cudaMemsetAsync(d_scal + 7, 0, 2 * sizeof(double), stream));
    allreduce(d_scal, 9);   // x-side sums are partial per x-block, y-side sums per row block
    if (timing) cudaEventRecord(te1, stream);
    fetch_scalars(9);
    HPR_CUDA_CHECK(cudaGetLastError());
    if (timing) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, te0, te1);
        g_rt.ms += ms; g_rt.wait_ms += (now_seconds() - host0) * 1e3; g_rt.n++;
        cudaEventDestroy(te0); cudaEventDestroy(te1);
        if (iter > 0 && iter % 100 == 0)
            fprintf(stderr, "[hprlp timing] residual checks so far: %d, device %.2f ms (%.3f ms each), host wait incl. queued iterations %.1f ms\n",
                    g_rt.n, g_rt.ms, g_rt.ms / g_rt.n, g_rt.wait_ms);
    }

    const double obj_scale = b_scale * c_scale;
    res->primal_obj = obj_scale * h_scal[1] + obj_constant;
    res->dual_obj = obj_scale * (h_scal[6] + h_scal[2]) + obj_constant;
    res->rel_gap = std::abs(res->primal_obj - res->dual_obj) / (1.0 + std::abs(res->primal_obj) + std::abs(res->dual_obj));
    res->err_Rd = c_scale * sqrt(h_scal[0]) / norm_c_org;
    res->err_Rp = b_scale * sqrt(h_scal[5]) / norm_b_org;
    if (iter == 0) res->err_Rp = std::max(res->err_Rp, b_scale * sqrt(h_scal[4]));
    res->kkt = std::max(std::max(res->err_Rd, res->err_Rp), res->rel_gap);

    if (compute_gap && rs != nullptr) {
        const double dot_prod = 2.0 * h_scal[7];
        const double dy2 = h_scal[8];
        const double dx2 = h_scal[3];
        double wnorm = sigma * (lambda_max * dy2) + dx2 / sigma + dot_prod;
        if (wnorm < 0) {
            printf("The estimated maximum eigenvalue is too small! Current value is %g\n", lambda_max);
            lambda_max = -(dot_prod + dx2 / sigma) / (sigma * dy2) * 1.05;
            printf("The new estimated maximum eigenvalue is %g\n", lambda_max);
            wnorm = sqrt(-(dot_prod + dx2 / sigma) * 0.05);
        } else {
            wnorm = sqrt(wnorm);
        }
        rs->current_gap = wnorm;
    }
}

// reference compute_weighted_norm, src/main_iterate.cu:486-515
double Engine::weighted_norm_after_restart() {
    const double *dxg = x_tmp;
    cudaTextureObject_t dxt = tex_xtmp;
    if (dist() && push_mode()) { exchange_x(false, 3, x_tmp); dxg = x_hat; dxt = tex_xhat; }   // A dx needs every block of dx
    else if (dist()) coll->all_gather_inplace(x_tmp, xblock, stream);
    WeightedNormOp o; o.dx = dxg; o.dy = y_tmp; o.tex = dxt; o.partials = d_partials;
    launch_stream(A, o, stream);
    final_reduce_kernel<<<1, 1024, 0, stream>>>(d_partials, part_blocks(A), 2, d_scal);
    sumsq_kernel<<<kVecBlocks, kVecThreads, 0, stream>>>(x_tmp + xb0, xb1 - xb0, d_partials);   // |dx|^2 over the owned block
    final_reduce_kernel<<<1, 1024, 0, stream>>>(d_partials, kVecBlocks, 1, d_scal + 2);
    allreduce(d_scal, 3);
    launches += 4;
    fetch_scalars(3);
    const double dot_prod = 2.0 * h_scal[0];
    const double dy2 = h_scal[1];
    const double dx2 = h_scal[2];
    double wnorm = sigma * (lambda_max * dy2) + dx2 / sigma + dot_prod;
    if (wnorm < 0) {
        printf("The estimated value of lambda_max is too small!\n");
        lambda_max = -(dot_prod + dx2 / sigma) / (sigma * dy2) * 1.05;
        wnorm = sqrt(-(dot_prod + dx2 / sigma) * 0.05);
    } else {
        wnorm = sqrt(wnorm);
    }
    return wnorm;
}

// reference update_sigma + do_restart + upload_halpern_restart_params,
// src/main_iterate.cu:367-404, 312-322, 54-66
void Engine::restart_and_sigma(RestartState *rs, const Residuals &res) {
    restart_kernel<<<kVecBlocks, kVecThreads, 0, stream>>>(x_bar, x0, x, xb0, xb1, y_bar, y0, y, m, d_partials);
    final_reduce_kernel<<<1, 1024, 0, stream>>>(d_partials, kVecBlocks, 2, d_scal);
    launches += 2;
    allreduce(d_scal, 2);   // |x_bar - x0|^2 is partial per x-block, |y_bar - y0|^2 per row block
    fetch_scalars(2);
    const double primal_move = sqrt(h_scal[0]);
    const double dual_move = sqrt(h_scal[1]);
    if (primal_move > 1e-16 && dual_move > 1e-16 && primal_move < 1e12 && dual_move < 1e12) {
        const double pm_over_dm = primal_move / dual_move;
        const double sqrt_lambda = sqrt(lambda_max);
        const double ratio = pm_over_dm / sqrt_lambda;
        const double fact = std::exp(-0.05 * (rs->current_gap / rs->best_gap));
        const double temp1 = std::max(std::min(res.err_Rd, res.err_Rp), std::min(res.rel_gap, rs->current_gap));
        const double sigma_cand = std::exp(fact * std::log(ratio) + (1 - fact) * std::log(rs->best_sigma));
        double kappa;
        if (temp1 > 9e-10) {
            kappa = 1.0;
        } else if (temp1 > 5e-10) {
            const double ratio_infeas = res.err_Rd / res.err_Rp;
            kappa = std::max(std::min(std::sqrt(ratio_infeas), 100.0), 1e-2);
        } else {
            const double ratio_infeas = res.err_Rd / res.err_Rp;
            kappa = std::max(std::min(ratio_infeas, 100.0), 1e-2);
        }
        sigma = kappa * sigma_cand;
    } else {
        sigma = 1.0;
    }
    rs->inner = 0;
    rs->times += 1;
    rs->save_gap = std::numeric_limits<double>::infinity();
    reset_halpern_counter();
}

void Engine::collect_solution(double *hx, double *hy, double *hz) {
    // unscale into scratch (wn, x_hat reused as z scratch is NOT allowed: x_hat is live) -> use wn/wm + x_tmp copy
    double *xo = wn, *yo = wm;
    double *zo = zo_buf;
    if (dist()) {   // every GPU owns one block of x_bar / z_bar
        coll->all_gather_inplace(x_bar, xblock, stream);
        coll->all_gather_inplace(z_bar, xblock, stream);
    }
    unscale_kernel<<<vec_grid(std::max(m, n)), kVecThreads, 0, stream>>>(x_bar, z_bar, col_norm, xo, zo, n, y_bar, row_norm, yo, m,
                                                                         b_scale, c_scale);
    launches++;
    HPR_CUDA_CHECK(cudaMemcpyAsync(hx, xo, sizeof(double) * n, cudaMemcpyDeviceToHost, stream));
    HPR_CUDA_CHECK(cudaMemcpyAsync(hz, zo, sizeof(double) * n, cudaMemcpyDeviceToHost, stream));
    if (!dist()) {
        HPR_CUDA_CHECK(cudaMemcpyAsync(hy, yo, sizeof(double) * m, cudaMemcpyDeviceToHost, stream));
        HPR_CUDA_CHECK(cudaStreamSynchronize(stream));
        return;
    }
    // y of the whole problem on every rank: the row blocks are placed in a zeroed m_global-vector and summed
    double *yfull = nullptr;
    HPR_CUDA_CHECK(cudaMalloc(&yfull, sizeof(double) * (size_t)m_global));
    HPR_CUDA_CHECK(cudaMemsetAsync(yfull, 0, sizeof(double) * (size_t)m_global, stream));
    HPR_CUDA_CHECK(cudaMemcpyAsync(yfull + row0, yo, sizeof(double) * m, cudaMemcpyDeviceToDevice, stream));
    allreduce(yfull, (size_t)m_global);
    HPR_CUDA_CHECK(cudaMemcpyAsync(hy, yfull, sizeof(double) * (size_t)m_global, cudaMemcpyDeviceToHost, stream));
    HPR_CUDA_CHECK(cudaStreamSynchronize(stream));
    cudaFree(yfull);
}

double Engine::time_phase_ms(int which, int reps) {
    cudaEvent_t e0, e1;
    HPR_CUDA_CHECK(cudaEventCreate(&e0));
    HPR_CUDA_CHECK(cudaEventCreate(&e1));
    auto one = [&]() {
        if (which == 0) launch_x_phase(false);
        else if (which == 2) partial_ATy_pass();   // row-partitioned x-side pass
        else if (which == 3)   // row-partitioned x-update on the owned block (from whatever wn holds)
            x_update_kernel<false><<<vec_grid(xb1 - xb0), kVecThreads, 0, stream>>>(wn, x, x_hat, c, l, u, x0, nullptr, nullptr, nullptr, d_params, d_k, d_k + 1, xb0, xb1);
        else if (which == 4) {   // the exchange of one iteration, as the loop issues it
            if (coll && push_mode()) exchange_x(false);   // peer-memory path: the exchange IS the fused x-update kernel
            else if (coll) { coll->reduce_scatter_inplace(wn, xblock, stream); coll->all_gather_inplace(x_hat, xblock, stream); }
        } else launch_y_phase(false);
        launches++;
    };
    one();
    HPR_CUDA_CHECK(cudaEventRecord(e0, stream));
    for (int i = 0; i < reps; ++i) one();
    HPR_CUDA_CHECK(cudaEventRecord(e1, stream));
    HPR_CUDA_CHECK(cudaEventSynchronize(e1));
    float ms = 0.f;
    HPR_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return (double)ms / reps;
}

// ------------------------------------------------------------------------------------------------
// the driver (reference HPRLP_main_solve from the start of the timed region, src/HPRLP.cu:150-311)
// ------------------------------------------------------------------------------------------------
static void check_restart(RestartState *r, int iter, int check_iter, double sigma) {
    // reference src/main_iterate.cu:324-364
    r->restart_flag = 0;
    if (r->first_restart) {
        if (iter == check_iter) {
            r->first_restart = false;
            r->restart_flag = 1;
            r->best_gap = r->current_gap;
            r->best_sigma = sigma;
        }
    } else if (iter % check_iter == 0) {
        if (r->current_gap < 0) {
            r->current_gap = 1e-6;
            printf("current_gap < 0\n");
        }
        if (r->current_gap <= 0.2 * r->last_gap) { r->sufficient += 1; r->restart_flag = 1; }
        if ((r->current_gap <= 0.6 * r->last_gap) && (r->current_gap > 1.00 * r->save_gap)) { r->necessary += 1; r->restart_flag = 2; }
        if (r->inner >= 0.2 * iter) { r->long_ += 1; r->restart_flag = 3; }
        if (r->best_gap > r->current_gap) { r->best_gap = r->current_gap; r->best_sigma = sigma; }
        r->save_gap = r->current_gap;
    }
}

void Engine::solve_begin(const HPRLP_parameters *param, SolveHooks *hooks) {
    LoopState &L = loop;
    L = LoopState();
    L.t_start_alg = now_seconds();
    const bool quiet = hooks->quiet;
    // lambda_max(A A^T) * 1.01 (reference compute_maximum_eigenvalue, src/HPRLP.cu:81-97)
    {
        const double t0 = now_seconds();
        int piters = 0;
        lambda_max = power_iteration(5000, 1e-4, hooks->power_z0, &piters) * 1.01;
        hooks->power_iters = piters;
        hooks->power_seconds = now_seconds() - t0;
        if (!quiet) printf("ESTIMATING MAXIMUM EIGENVALUE time = %.2f seconds\n", hooks->power_seconds);
    }
    sigma = (norm_b > 1e-8 && norm_c > 1e-8) ? norm_b / norm_c : 1.0;
    L.rs.best_sigma = sigma;
    init_iterates();
    L.output.residuals = 0; L.output.primal_obj = 0; L.output.gap = 0;
    std::memset(L.output.status, 0, sizeof(L.output.status));
    if (!quiet) {
        printf(" iter     errRp        errRd         p_obj            d_obj          gap         sigma       time\n");
        fflush(stdout);
    }
    L.uploaded_sigma = sigma;
    L.uploaded_lambda = lambda_max;
    L.iter = 0;
    (void)param;
}

// Advances the driver.  Returns true when a stopping status was reached (L.output is final), false
// when paused at iteration index `pause_at` (>0; bench/step-wise use; pause indices should be
// multiples of step() so that they coincide with the reference's own residual iterations).
bool Engine::solve_advance(const HPRLP_parameters *param, SolveHooks *hooks, int pause_at) {
    LoopState &L = loop;
    RestartState &rs = L.rs;
    Residuals &res = L.res;
    HPRLP_results &output = L.output;
    const bool quiet = hooks->quiet;
    const double t_start_alg = L.t_start_alg;
    const int check_iter = std::max(param->check_iter, 1);
    bool first_pass = true;
    for (;;) {
        int &iter = L.iter;
        if (pause_at > 0 && iter >= pause_at && !first_pass) return false;
        first_pass = false;
        // every index visited here is "eventful": periodic, print or max_iter (see next_event below)
        const bool periodic = (iter % check_iter == 0);
        const bool compute_gap = periodic && iter > 0;
        compute_residuals(iter, compute_gap, &res, &rs);

        double elapsed = now_seconds() - t_start_alg;
        if (dist()) {   // every rank must take the same TIME_LIMIT decision: use the maximum over ranks
            h_scal[15] = elapsed;
            HPR_CUDA_CHECK(cudaMemcpyAsync(d_scal + 15, h_scal + 15, sizeof(double), cudaMemcpyHostToDevice, stream));
            allreduce(d_scal + 15, 1, true);
            HPR_CUDA_CHECK(cudaMemcpyAsync(h_scal + 15, d_scal + 15, sizeof(double), cudaMemcpyDeviceToHost, stream));
            HPR_CUDA_CHECK(cudaStreamSynchronize(stream));
            elapsed = h_scal[15];
        }
        const char *status = "CONTINUE";   // reference check_stopping, src/main_iterate.cu:406-420
        if (res.kkt < param->stop_tol) status = "OPTIMAL";
        else if (iter >= param->max_iter) status = "ITER_LIMIT";
        else if (elapsed > param->time_limit) status = "TIME_LIMIT";

        if (periodic) check_restart(&rs, iter, check_iter, sigma);
        else rs.restart_flag = 0;

        // reference print_flag, src/HPRLP.cu:184-186,207
        const bool print_flag = (iter % step_of(iter) == 0) || (iter == param->max_iter) || (elapsed > param->time_limit);
        if (!quiet && (print_flag || std::strcmp(status, "CONTINUE") != 0)) {
            printf("%5d    %.2e    %.2e    %+.6e    %+.6e    %.2e    %.2e      %.2f\n", iter, res.err_Rp, res.err_Rd,
                   res.primal_obj, res.dual_obj, res.rel_gap, sigma, now_seconds() - t_start_alg);
            fflush(stdout);
        }
        if (L.first_4 && res.kkt < 1e-4) {
            output.iter4 = iter; output.time4 = now_seconds() - t_start_alg; L.first_4 = false;
            if (!quiet) printf("Residual < 1e-4 at iter = %d\n", iter);
        }
        if (L.first_6 && res.kkt < 1e-6) {
            output.iter6 = iter; output.time6 = now_seconds() - t_start_alg; L.first_6 = false;
            if (!quiet) printf("Residual < 1e-6 at iter = %d\n", iter);
        }
        if (L.first_8 && res.kkt < 1e-8) {
            output.iter8 = iter; output.time8 = now_seconds() - t_start_alg; L.first_8 = false;
            if (!quiet) printf("Residual < 1e-8 at iter = %d\n", iter);
        }

        if (std::strcmp(status, "CONTINUE") != 0) {
            std::strncpy(output.status, status, sizeof(output.status) - 1);
            output.iter = iter;
            output.gap = res.rel_gap;
            output.residuals = res.kkt;
            output.primal_obj = res.primal_obj;
            output.time = now_seconds() - t_start_alg;
            output.time4 = (output.time4 == 0.0) ? output.time : output.time4;
            output.time6 = (output.time6 == 0.0) ? output.time : output.time6;
            output.time8 = (output.time8 == 0.0) ? output.time : output.time8;
            output.iter4 = (output.iter4 == 0) ? output.iter : output.iter4;
            output.iter6 = (output.iter6 == 0) ? output.iter : output.iter6;
            output.iter8 = (output.iter8 == 0) ? output.iter : output.iter8;
            output.x = static_cast<double *>(std::malloc(sizeof(double) * n));
            output.y = static_cast<double *>(std::malloc(sizeof(double) * (size_t)mg()));
            output.z = static_cast<double *>(std::malloc(sizeof(double) * n));
            collect_solution(output.x, output.y, output.z);
            if (!quiet) {
                printf("\n=== Solution Summary ===\nStatus: %s\nIterations: %d\nTime: %.2f seconds\n", output.status, output.iter,
                       output.time);
                printf("Primal Objective: %.12e\nResidual: %.12e\n\n", output.primal_obj, output.residuals);
                fflush(stdout);
            }
            return true;
        }

        const bool restart = rs.restart_flag > 0;
        if (restart) restart_and_sigma(&rs, res);
        if (sigma != L.uploaded_sigma || lambda_max != L.uploaded_lambda) {
            upload_params();
            L.uploaded_sigma = sigma;
            L.uploaded_lambda = lambda_max;
        }

        // next eventful index: next multiple of check_iter, next multiple of step(), max_iter, or the pause
        int next_event;
        {
            const long long nper = ((long long)iter / check_iter + 1) * check_iter;
            const long long i1 = (long long)iter + 1;
            const long long st = step_of((int)std::min<long long>(i1, INT32_MAX));
            const long long nprint = ((i1 + st - 1) / st) * st;
            long long ne = std::min(nper, nprint);
            if ((long long)param->max_iter > iter) ne = std::min(ne, (long long)param->max_iter);
            if (pause_at > iter) ne = std::min(ne, (long long)pause_at);
            // TIME_LIMIT: the reference looks at the clock every iteration (src/HPRLP.cu:184-198).  Launches are asynchronous
            // here, so the run to the next visited index is capped by what the remaining time allows at the rate measured
            // so far; at that index the clock is read again (after the residual fetch has synchronised).
            if (iter > 0 && param->time_limit < 1e9) {
                const double per_iter = elapsed / iter;
                const double left = param->time_limit - elapsed;
                const long long fit = per_iter > 0 ? (long long)std::max(1.0, std::min(left / per_iter + 1.0, 2.0e9)) : 1;
                ne = std::min(ne, (long long)iter + std::max<long long>(fit, 1));
            }
            next_event = (int)std::min<long long>(ne, INT32_MAX);
        }
        // iterations iter .. next_event-1: the last one is a check iteration when the reference's
        // (iter+1)%check_iter==0 || (iter+1)%step(iter+1)==0 holds (src/HPRLP.cu:295-296) -- true for every
        // eventful index except a max_iter that is not a multiple of step (quirk: stale bars at ITER_LIMIT).
        const int count = next_event - iter;
        const bool last_is_check = (next_event % check_iter == 0) || (next_event % step_of(next_event) == 0);
        int done = 0;
        if (restart) {
            launch_iteration(true);
            done = 1;
            rs.last_gap = weighted_norm_after_restart();
            if (lambda_max != L.uploaded_lambda) { upload_params(); L.uploaded_lambda = lambda_max; }
        }
        const int tail_check = (last_is_check && count - done >= 1) ? 1 : 0;
        run_normal(count - done - tail_check);
        if (tail_check) launch_iteration(true);
        rs.inner += count;
        iter = next_event;

        for (int t = 0; t < hooks->n_trace; ++t) {
            if (hooks->trace_iters[t] == iter && (last_is_check || (restart && count == 1))) {
                collect_solution(hooks->trace_x + (size_t)t * n, hooks->trace_y + (size_t)t * mg(), hooks->trace_z + (size_t)t * n);
            }
        }
    }
}

void Engine::fill_hooks(SolveHooks *hooks) {
    hooks->lambda_max = lambda_max;
    hooks->sigma = sigma;
    hooks->restarts = loop.rs.times;
    hooks->kernel_launches = launches;
    hooks->scal[0] = b_scale; hooks->scal[1] = c_scale; hooks->scal[2] = norm_b; hooks->scal[3] = norm_c;
    hooks->scal[4] = norm_b_org; hooks->scal[5] = norm_c_org;
}

HPRLP_results Engine::solve(const HPRLP_parameters *param, SolveHooks *hooks) {
    SolveHooks local;
    if (!hooks) hooks = &local;
    cudaEvent_t ev0, ev1;
    HPR_CUDA_CHECK(cudaEventCreate(&ev0));
    HPR_CUDA_CHECK(cudaEventCreate(&ev1));
    solve_begin(param, hooks);
    HPR_CUDA_CHECK(cudaEventRecord(ev0, stream));
    solve_advance(param, hooks, -1);
    HPR_CUDA_CHECK(cudaEventRecord(ev1, stream));
    HPR_CUDA_CHECK(cudaEventSynchronize(ev1));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev0, ev1);
    hooks->loop_device_ms = ms;
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    fill_hooks(hooks);
    return loop.output;
}

}  // namespace hpr
