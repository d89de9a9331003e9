// synth_device.cu -- on-device generation of the synthetic "uniform" LP of BASELINE.json for the
// row-partitioned multi-GPU path: instances whose CSR + transpose exceed one GPU (config 5, nnz = 6e9) cannot be
// built on the host through the int32 C ABI, so every GPU generates its own row block.  Same counter-based RNG and
// the same per-row procedure as tools/synth_lp.c (splitmix64 of (seed, stream, index); K distinct sorted columns per
// row with duplicates re-drawn; values U(-1,1) with |v| >= 1e-3), so a block generated here is bit-identical to the
// host generator's rows (tests/test_gpu_partitioned.py checks it).  Test/bench infrastructure that lives in the
// library only because it must write straight into engine buffers.
#include <chrono>
#include <cmath>
#include <memory>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/HPRLP.h"
#include "../../include/hprlp_b200.h"
#include "abi_guard.h"
#include "engine.h"
#include "rank_group.h"

namespace hpr {
namespace {

__host__ __device__ inline unsigned long long sm64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
__host__ __device__ inline unsigned long long rng(unsigned long long seed, unsigned long long stream, unsigned long long idx) {
    return sm64(sm64(seed ^ (stream * 0xD6E8FEB86659FD93ULL)) + idx * 0x9E3779B97F4A7C15ULL);
}
__host__ __device__ inline double u01(unsigned long long r) { return (double)(r >> 11) * (1.0 / 9007199254740992.0); }

constexpr int kGenWarps = 4;

// one warp per row: K distinct sorted columns (bitonic sort in shared memory, sequential re-draw of duplicates exactly
// as the host generator does), then K values by ordered rejection sampling
__global__ void gen_uniform_rows_kernel(int rows_local, long long row0, int K, int KP, int n, unsigned long long seed,
                                        int *rowPtr, int *col, double *val) {
    extern __shared__ int csm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.x * kGenWarps + warp;
    if (i >= rows_local) return;
    int *c = csm + (size_t)warp * KP;
    const unsigned long long gi = (unsigned long long)(row0 + i);
    const unsigned long long sc = 2ULL + 4ULL * gi, sv = 3ULL + 4ULL * gi;
    for (int k = lane; k < KP; k += 32) c[k] = k < K ? (int)(rng(seed, sc, (unsigned long long)k) % (unsigned long long)n) : 0x7fffffff;
    unsigned long long ctr = (unsigned long long)K;
    for (;;) {
        __syncwarp();
        for (int size = 2; size <= KP; size <<= 1)
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                for (int t = lane; t < KP / 2; t += 32) {
                    const int a = 2 * t - (t & (stride - 1)), b = a + stride;
                    const bool up = (a & size) == 0;
                    const int va = c[a], vb = c[b];
                    if ((va > vb) == up) { c[a] = vb; c[b] = va; }
                }
                __syncwarp();
            }
        int dup = 0;
        for (int k = 1 + lane; k < K; k += 32) dup |= (c[k] == c[k - 1]);
        if (!__any_sync(0xffffffffu, dup)) break;
        if (lane == 0)
            for (int k = 1; k < K; ++k)
                if (c[k] == c[k - 1]) c[k - 1] = (int)(rng(seed, sc, ctr++) % (unsigned long long)n);
    }
    const size_t base = (size_t)i * K;
    for (int k = lane; k < K; k += 32) col[base + k] = c[k];
    int out = 0;
    for (unsigned long long j0 = 0; out < K; j0 += 32) {
        const double v = 2.0 * u01(rng(seed, sv, j0 + lane)) - 1.0;
        const bool acc = fabs(v) >= 1e-3;
        const unsigned ball = __ballot_sync(0xffffffffu, acc);
        const int pos = out + __popc(ball & ((1u << lane) - 1u));
        if (acc && pos < K) val[base + pos] = v;
        out += __popc(ball);
    }
    if (lane == 0) {
        rowPtr[i] = (int)((size_t)i * K);
        if (i == rows_local - 1) rowPtr[rows_local] = (int)((size_t)rows_local * K);
    }
}

// x*, z*, l, u per column (tools/synth_lp.c synth_lp_vectors, column loop)
__global__ void gen_col_vectors_kernel(int n, unsigned long long seed, unsigned long long S, double *l, double *u, double *xs, double *zs) {
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const bool at_zero = (rng(S, 11, (unsigned long long)j) & 1ULL) != 0;
        l[j] = 0.0;
        u[j] = (rng(seed, 12, (unsigned long long)j) % 10ULL == 0) ? 1.0 : INFINITY;
        if (at_zero) { xs[j] = 0.0; zs[j] = u01(rng(S, 13, (unsigned long long)j)); }
        else { xs[j] = 0.05 + 0.9 * u01(rng(S, 14, (unsigned long long)j)); zs[j] = 0.0; }
    }
}
// AL, AU, y* per row from ax = (A x*)_i (row loop of synth_lp_vectors)
__global__ void gen_row_vectors_kernel(int m_local, long long row0, unsigned long long seed, unsigned long long S, const double *ax,
                                       double *AL, double *AU, double *ys) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m_local; i += gridDim.x * blockDim.x) {
        const unsigned long long gi = (unsigned long long)(row0 + i);
        const unsigned long long type = rng(seed, 15, gi) % 3ULL;
        const bool active = (rng(seed, 16, gi) & 1ULL) != 0;
        const double r1 = u01(rng(S, 17, gi)), r2 = u01(rng(S, 18, gi)), a = ax[i];
        if (type == 0) { AL[i] = a; AU[i] = a; ys[i] = 2.0 * r1 - 1.0; }
        else if (type == 1) {
            AL[i] = -INFINITY;
            if (active) { AU[i] = a; ys[i] = -r1; } else { AU[i] = a + 0.1 + r2; ys[i] = 0.0; }
        } else {
            if (active) { AL[i] = a; AU[i] = a + 0.5 + r2; ys[i] = r1; }
            else { AL[i] = a - 0.1 - r1; AU[i] = a + 0.1 + r2; ys[i] = 0.0; }
        }
    }
}
// c = A'y* (already reduced over the row blocks) + z*; obj += <c, x*>
__global__ void gen_cost_kernel(int n, const double *w, const double *zs, const double *xs, double *c, double *obj) {
    double t = 0.0;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const double cj = w[j] + zs[j];
        c[j] = cj;
        t += cj * xs[j];
    }
    for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
    if ((threadIdx.x & 31) == 0) atomicAdd(obj, t);
}

int pow2_at_least(int k) { int p = 32; while (p < k) p <<= 1; return p; }

void generate_rows(int rows_local, long long row0, int K, int n, unsigned long long seed, int *rowPtr, int *col, double *val, cudaStream_t st) {
    const int KP = pow2_at_least(K);
    const size_t smem = sizeof(int) * (size_t)kGenWarps * KP;
    gen_uniform_rows_kernel<<<(rows_local + kGenWarps - 1) / kGenWarps, 32 * kGenWarps, smem, st>>>(rows_local, row0, K, KP, n, seed, rowPtr, col, val);
}

}  // namespace
}  // namespace hpr

using namespace hpr;

// Debug/test hook: rows [row0, row0+rows) of the synthetic uniform matrix generated on the device, copied to the host.
extern "C" int hprlp_b200_synth_rows(int n, int K, unsigned long long seed, long long row0, int rows, int *col_out, double *val_out) {
    return abi_guard_int("hprlp_b200_synth_rows", [&]() -> int {
        if (rows <= 0 || K <= 0 || K > 4096 || K > n) return -1;
        int *rp = nullptr, *col = nullptr;
        double *val = nullptr;
        HPR_CUDA_CHECK(cudaMalloc(&rp, sizeof(int) * ((size_t)rows + 1)));
        HPR_CUDA_CHECK(cudaMalloc(&col, sizeof(int) * (size_t)rows * K));
        HPR_CUDA_CHECK(cudaMalloc(&val, sizeof(double) * (size_t)rows * K));
        generate_rows(rows, row0, K, n, seed, rp, col, val, 0);
        HPR_CUDA_CHECK(cudaMemcpy(col_out, col, sizeof(int) * (size_t)rows * K, cudaMemcpyDeviceToHost));
        HPR_CUDA_CHECK(cudaMemcpy(val_out, val, sizeof(double) * (size_t)rows * K, cudaMemcpyDeviceToHost));
        cudaFree(rp); cudaFree(col); cudaFree(val);
        return 0;
    });
}

namespace {

struct SynthSpec { long long m; int n, K; unsigned long long seed; };

void check_synth(const SynthSpec &s, int P) {
    if (s.m <= 0 || s.n <= 0 || s.K <= 0 || s.K > s.n || s.K > 4096 || s.m > 2147483647LL) throw std::runtime_error("synth: bad dimensions");
    if ((s.m + P - 1) / P * (long long)s.K >= 2147483647LL) throw std::runtime_error("synth: a shard would exceed 2^31 nonzeros; use more GPUs");
}

// One rank: generate rows [m p/P, m (p+1)/P) and the vectors on the device, scale, solve.  coll == nullptr: single GPU.
HPRLP_results synth_rank(const SynthSpec &sp, int p, int P, int device, Collective *coll, const HPRLP_parameters &param, bool quiet,
                         double *obj_out, hprlp_b200_info *info) {
    const long long r0 = sp.m * p / P, r1 = sp.m * (p + 1) / P;
    const int mp = (int)(r1 - r0), n = sp.n, K = sp.K;
    const long long nnzp = (long long)mp * K;
    HPRLP_parameters pp = param;
    pp.device_number = device;
    Engine eng;
    if (coll) eng.set_partition(coll, (int)sp.m, (int)r0);
    SolveHooks hooks;
    hooks.quiet = quiet;
    const auto t0 = std::chrono::steady_clock::now();
    eng.prepare(mp, n, nnzp, device);
    generate_rows(mp, r0, K, n, sp.seed, eng.A.rowPtr, eng.A.col, eng.A.val, eng.stream);
    eng.finish_setup(true);
    // vectors: x*, z* live in x_bar / z_bar until init_iterates() clears them
    double *xs = eng.x_bar, *zs = eng.z_bar, *ys = eng.wm2, *ax = eng.wm, *w = eng.wn;
    gen_col_vectors_kernel<<<1184, 256, 0, eng.stream>>>(n, sp.seed, sp.seed, eng.l, eng.u, xs, zs);
    eng.spmv_A(xs, ax);
    gen_row_vectors_kernel<<<1184, 256, 0, eng.stream>>>(mp, r0, sp.seed, sp.seed, ax, eng.AL, eng.AU, ys);
    eng.spmv_AT(ys, w);
    eng.allreduce(w, (size_t)n);
    HPR_CUDA_CHECK(cudaMemsetAsync(eng.d_scal, 0, sizeof(double), eng.stream));
    gen_cost_kernel<<<1184, 256, 0, eng.stream>>>(n, w, zs, xs, eng.c, eng.d_scal);
    double obj = 0.0;
    HPR_CUDA_CHECK(cudaMemcpyAsync(&obj, eng.d_scal, sizeof(double), cudaMemcpyDeviceToHost, eng.stream));
    HPR_CUDA_CHECK(cudaStreamSynchronize(eng.stream));
    if (obj_out) *obj_out = obj;
    hooks.setup_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    const auto t1 = std::chrono::steady_clock::now();
    eng.scale(&pp);
    hooks.scaling_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count();
    HPRLP_results r = eng.solve(&pp, &hooks);
    fill_b200_info(eng, hooks, info);
    return r;
}

}  // namespace

// Row-partitioned solve of the synthetic uniform LP (m rows, n columns, K nonzeros per row) generated shard by shard
// on the GPUs: the path for BASELINE.json config 5 (nnz = 6e9).  Each shard must have < 2^31 nonzeros.
// obj_star receives the constructed optimal value <c, x*>.  With want_solution == 0 the x/y/z vectors are not
// returned (result.x/y/z = NULL).  One host thread per GPU.
extern "C" HPRLP_results hprlp_b200_solve_partitioned_synth(long long m, int n, int K, unsigned long long seed,
                                                             const HPRLP_parameters *param_in, int n_gpus, int quiet,
                                                             int want_solution, double *obj_star, hprlp_b200_info *info) {
    return abi_guard_results("hprlp_b200_solve_partitioned_synth", [&]() -> HPRLP_results {
        HPRLP_parameters def;
        const HPRLP_parameters param = param_in ? *param_in : def;
        int avail = 0;
        if (cudaGetDeviceCount(&avail) != cudaSuccess || avail < 1) throw std::runtime_error("no CUDA device");
        const int P = std::max(1, std::min(n_gpus, avail - param.device_number));
        const SynthSpec sp{m, n, K, seed};
        check_synth(sp, P);
        if (!quiet) std::printf("Synthetic uniform LP m=%lld n=%d nnz=%lld generated on %d GPU(s), row-block partitioned\n", m, n, m * K, P);
        HPRLP_results out = abi_error_result();
        hprlp_b200_info info0{};
        double obj0 = 0.0;
        if (P == 1) {
            out = synth_rank(sp, 0, 1, param.device_number, nullptr, param, quiet != 0, &obj0, &info0);
        } else {
            std::vector<int> devs(P);
            for (int p = 0; p < P; ++p) devs[p] = param.device_number + p;
            std::vector<HPRLP_results> results(P, abi_error_result());
            RankGroup group(Transport::Nccl, devs);
            try {
                group.run([&](int p, int device, Collective *coll) {
                    results[p] = synth_rank(sp, p, P, device, coll, param, quiet != 0 || p != 0, p == 0 ? &obj0 : nullptr, p == 0 ? &info0 : nullptr);
                });
            } catch (...) {
                for (auto &r : results) { std::free(r.x); std::free(r.y); std::free(r.z); }
                throw;
            }
            for (int p = 1; p < P; ++p) { std::free(results[p].x); std::free(results[p].y); std::free(results[p].z); }
            out = results[0];
        }
        if (!want_solution) { std::free(out.x); std::free(out.y); std::free(out.z); out.x = out.y = out.z = nullptr; }
        if (obj_star) *obj_star = obj0;
        if (info) *info = info0;
        return out;
    });
}

// The same, one process per GPU: this process generates and owns row block `rank` of `nranks` (NCCL communicator from
// the unique id the caller distributed, see hprlp_b200_nccl_unique_id).
extern "C" HPRLP_results hprlp_b200_solve_partitioned_synth_rank(long long m, int n, int K, unsigned long long seed,
                                                                  const HPRLP_parameters *param_in, hprlp_b200_comm *comm,
                                                                  int quiet, int want_solution, double *obj_star,
                                                                  hprlp_b200_info *info) {
    return abi_guard_results("hprlp_b200_solve_partitioned_synth_rank", [&]() -> HPRLP_results {
        HPRLP_parameters def;
        const HPRLP_parameters param = param_in ? *param_in : def;
        Collective *coll = comm ? comm->coll.get() : nullptr;
        const int nranks = coll ? coll->nranks : 1, rank = coll ? coll->rank : 0;
        const SynthSpec sp{m, n, K, seed};
        check_synth(sp, nranks);
        HPRLP_results out = synth_rank(sp, rank, nranks, comm ? comm->device : param.device_number, coll, param, quiet != 0, obj_star, info);
        if (!want_solution) { std::free(out.x); std::free(out.y); std::free(out.z); out.x = out.y = out.z = nullptr; }
        return out;
    });
}
