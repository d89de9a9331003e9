// partitioned.cu -- one LP row-block partitioned over the GPUs of one node (new functionality: the
// reference is single-GPU, src/HPRLP.cu:51-64).  GPU p owns
//   * a contiguous block of rows of A (balanced by nonzeros) as CSR A_p and its transpose A_p^T (built on the device),
//     with the matching blocks of y / AL / AU / row_norm, and
//   * the x-block J_p (n/P columns): x, x0, x_bar, z_bar live there only.
// Per HPR iteration (Engine::launch_iteration):
//   partial w_p = A_p^T y_p  ->  reduce-scatter (sum over p, GPU p keeps w on J_p)  ->  x-update on J_p
//   ->  all-gather of the x_hat blocks  ->  fused SpMV + projection + Halpern y-phase on the local rows.
// Checks: the <= 9 residual / restart scalars are all-reduced; the primal residual pass needs one all-gather of x_bar.
// Setup: column statistics of the scaling and the power iteration's A^T q are all-reduced n-vectors.
// Every rank runs the same deterministic host logic on identical reduced scalars.
//
// Entry points: threads (one process drives all GPUs, ncclCommInitAll), ranks (one process per GPU, ncclCommInitRank
// with a unique id the caller distributes -- bench.py does it with torch.distributed under torchrun), and "local"
// (P logical ranks on ONE GPU with host-synchronised exchanges: the parity-test path for 1-GPU boxes).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/HPRLP.h"
#include "../../include/hprlp_b200.h"
#include "abi_guard.h"
#include "engine.h"
#include "rank_group.h"

namespace hpr {

// row blocks balanced by nonzeros: block p = rows [b[p], b[p+1])
std::vector<int> row_blocks_by_nnz(const int *rp, int m, int P) {
    const long long nnz = rp[m];
    std::vector<int> b(P + 1, 0);
    b[P] = m;
    for (int p = 1; p < P; ++p) {
        const long long target = nnz * p / P;
        b[p] = (int)(std::lower_bound(rp, rp + m + 1, (int)target) - rp);
        b[p] = std::max(b[p], b[p - 1]);
        b[p] = std::min(b[p], m);
    }
    return b;
}

// rows [r0, r1) of `model` + the (replicated) column vectors onto eng's device; eng.set_partition() comes first
void upload_row_block(Engine &eng, const LP_info_cpu *model, int r0, int r1, int device) {
    const int *rp = model->A->rowPtr;
    const int mp = r1 - r0;
    std::vector<int> rp_local((size_t)mp + 1);
    for (int i = 0; i <= mp; ++i) rp_local[i] = rp[r0 + i] - rp[r0];
    sparseMatrix Ap{mp, model->n, rp_local[mp], model->A->colIndex + rp[r0], rp_local.data(), model->A->value + rp[r0]};
    LP_info_cpu shard{};
    shard.m = mp; shard.n = model->n; shard.A = &Ap;
    shard.AL = model->AL + r0; shard.AU = model->AU + r0;
    shard.c = model->c; shard.l = model->l; shard.u = model->u;
    shard.obj_constant = model->obj_constant;
    eng.upload(&shard, device);
}

Collective *open_nccl_rank(const char *uid128, int rank, int nranks, int device) {
    HPR_CUDA_CHECK(cudaSetDevice(device));
    NcclUniqueId id;
    std::memcpy(&id, uid128, sizeof(id));
    NcclComm comm = nullptr;
    const int rc = nccl().CommInitRank(&comm, nranks, id, rank);
    if (rc != 0) throw std::runtime_error(std::string("ncclCommInitRank failed: ") + nccl().GetErrorString(rc));
    return make_nccl_collective(comm, nranks, rank);
}

void fill_b200_info(const Engine &eng, const SolveHooks &h, hprlp_b200_info *info) {
    if (!info) return;
    info->lambda_max = h.lambda_max; info->sigma = h.sigma;
    info->setup_seconds = h.setup_seconds; info->scaling_seconds = h.scaling_seconds; info->power_seconds = h.power_seconds;
    info->loop_device_ms = h.loop_device_ms;
    info->restarts = h.restarts; info->power_iters = h.power_iters; info->kernel_launches = h.kernel_launches;
    info->b_scale = h.scal[0]; info->c_scale = h.scal[1]; info->norm_b = h.scal[2]; info->norm_c = h.scal[3];
    info->norm_b_org = h.scal[4]; info->norm_c_org = h.scal[5];
    info->lanes_A = eng.A.G; info->lanes_AT = eng.AT.G; info->items_A = eng.A.n_items; info->items_AT = eng.AT.n_items;
    info->bands_A = (int)eng.A.bands.size(); info->reserved0 = eng.nranks;
    info->peer_exchange = eng.px ? 1 : 0; info->reserved1 = 0;
}

namespace {

void free_xyz(HPRLP_results &r) {
    std::free(r.x); std::free(r.y); std::free(r.z);
    r.x = r.y = r.z = nullptr;
}

// one rank of a partitioned solve of a host model: upload the row block, scale, solve
HPRLP_results solve_rank(const LP_info_cpu *model, const std::vector<int> &b, int p, int device, Collective *coll,
                         const HPRLP_parameters &param, bool quiet, hprlp_b200_info *info) {
    HPRLP_parameters pp = param;
    pp.device_number = device;
    if (const char *e = std::getenv("HPRLP_TEST_FAIL_RANK"))   // tests: this rank fails before its first collective
        if (std::atoi(e) == p) throw std::runtime_error("injected failure (HPRLP_TEST_FAIL_RANK)");
    Engine eng;
    eng.set_partition(coll, model->m, b[p]);
    SolveHooks hooks;
    hooks.quiet = quiet;
    upload_row_block(eng, model, b[p], b[p + 1], device);
    eng.scale(&pp);
    HPRLP_results r = eng.solve(&pp, &hooks);
    fill_b200_info(eng, hooks, info);
    return r;
}

HPRLP_results solve_threads(const LP_info_cpu *model, const HPRLP_parameters &param, Transport transport, int P,
                            const std::vector<int> &devs, int quiet, hprlp_b200_info *info) {
    const std::vector<int> b = row_blocks_by_nnz(model->A->rowPtr, model->m, P);
    if (!quiet) {
        std::printf("Row-block partition over %d %s (reduce-scatter of A^T y, all-gather of x_hat per iteration): rows", P,
                    transport == Transport::Nccl ? "GPUs, NCCL" : "logical ranks on one GPU");
        for (int p = 0; p <= P; ++p) std::printf(" %d", b[p]);
        std::printf("\n");
    }
    std::vector<HPRLP_results> results(P, abi_error_result());
    hprlp_b200_info info0{};
    RankGroup group(transport, devs);
    try {
        group.run([&](int p, int device, Collective *coll) {
            results[p] = solve_rank(model, b, p, device, coll, param, quiet != 0 || p != 0, p == 0 ? &info0 : nullptr);
        });
    } catch (...) {
        for (auto &r : results) free_xyz(r);
        throw;
    }
    for (int p = 1; p < P; ++p) free_xyz(results[p]);   // every rank returns the full solution; rank 0's is handed out
    if (info) *info = info0;
    return results[0];
}

}  // namespace
}  // namespace hpr

using namespace hpr;

extern "C" HPRLP_results hprlp_b200_solve_partitioned(const LP_info_cpu *model, const HPRLP_parameters *param_in, int n_gpus,
                                                       int quiet, hprlp_b200_info *info) {
    return abi_guard_results("hprlp_b200_solve_partitioned", [&]() -> HPRLP_results {
        HPRLP_parameters def;
        const HPRLP_parameters param = param_in ? *param_in : def;
        if (!model || !model->A) { abi_report("hprlp_b200_solve_partitioned", "Null model pointer"); return abi_error_result(); }
        int avail = 0;
        if (cudaGetDeviceCount(&avail) != cudaSuccess || avail < 1) throw std::runtime_error("no CUDA device");
        const int P = std::max(1, std::min(std::min(n_gpus, avail - param.device_number), model->m));
        if (P == 1) return hprlp_b200_solve_ex(model, &param, nullptr, 0, nullptr, nullptr, nullptr, nullptr, quiet, info);
        std::vector<int> devs(P);
        for (int p = 0; p < P; ++p) devs[p] = param.device_number + p;
        return solve_threads(model, param, Transport::Nccl, P, devs, quiet, info);
    });
}

// P logical ranks on ONE GPU (param->device_number): the same partitioned engine code, exchanges done by plain kernels
// between host barriers.  For parity tests on single-GPU boxes, not for speed.
extern "C" HPRLP_results hprlp_b200_solve_partitioned_local(const LP_info_cpu *model, const HPRLP_parameters *param_in,
                                                             int n_ranks, int quiet, hprlp_b200_info *info) {
    return abi_guard_results("hprlp_b200_solve_partitioned_local", [&]() -> HPRLP_results {
        HPRLP_parameters def;
        const HPRLP_parameters param = param_in ? *param_in : def;
        if (!model || !model->A) { abi_report("hprlp_b200_solve_partitioned_local", "Null model pointer"); return abi_error_result(); }
        const int P = std::max(1, std::min(std::min(n_ranks, 16), model->m));
        std::vector<int> devs(P, param.device_number);
        return solve_threads(model, param, Transport::Local, P, devs, quiet, info);
    });
}

// ---- one process per GPU -------------------------------------------------------------------------------------------
extern "C" int hprlp_b200_nccl_unique_id(char *out128) {
    return abi_guard_int("hprlp_b200_nccl_unique_id", [&]() -> int {
        if (!out128) return -1;
        NcclUniqueId id;
        const int rc = nccl().GetUniqueId(&id);
        if (rc != 0) throw std::runtime_error(std::string("ncclGetUniqueId failed: ") + nccl().GetErrorString(rc));
        std::memcpy(out128, &id, sizeof(id));
        return 0;
    });
}

extern "C" hprlp_b200_comm *hprlp_b200_comm_create(const char *uid128, int rank, int nranks, int device) {
    return abi_guard<hprlp_b200_comm *>("hprlp_b200_comm_create", [&]() -> hprlp_b200_comm * {
        if (!uid128 || nranks < 2 || rank < 0 || rank >= nranks) throw std::runtime_error("bad arguments");
        std::unique_ptr<hprlp_b200_comm> c(new hprlp_b200_comm);
        c->device = device;
        c->coll.reset(open_nccl_rank(uid128, rank, nranks, device));
        return c.release();
    }, [] { return (hprlp_b200_comm *)nullptr; });
}
extern "C" void hprlp_b200_comm_destroy(hprlp_b200_comm *comm) { delete comm; }

extern "C" HPRLP_results hprlp_b200_solve_partitioned_rank(const LP_info_cpu *model, const HPRLP_parameters *param_in,
                                                            hprlp_b200_comm *comm, int quiet, hprlp_b200_info *info) {
    return abi_guard_results("hprlp_b200_solve_partitioned_rank", [&]() -> HPRLP_results {
        HPRLP_parameters def;
        const HPRLP_parameters param = param_in ? *param_in : def;
        if (!model || !model->A || !comm || !comm->coll || comm->coll->nranks > model->m) {
            abi_report("hprlp_b200_solve_partitioned_rank", "bad arguments");
            return abi_error_result();
        }
        Collective *coll = comm->coll.get();
        const std::vector<int> b = row_blocks_by_nnz(model->A->rowPtr, model->m, coll->nranks);
        return solve_rank(model, b, coll->rank, comm->device, coll, param, quiet != 0, info);
    });
}

// Diagnostic: average time (ms) of the per-iteration exchange pair -- in-place reduce-scatter + all-gather of `count`
// doubles -- and of one all-reduce of the same vector, issued from one host thread per GPU exactly as the partitioned
// solver issues them.  out_ms[0] = reduce-scatter + all-gather, out_ms[1] = all-reduce.
extern "C" int hprlp_b200_nccl_exchange_ms(int n_gpus, long long count, int reps, double *out_ms) {
    return abi_guard_int("hprlp_b200_nccl_exchange_ms", [&]() -> int {
        int avail = 0;
        cudaGetDeviceCount(&avail);
        const int P = std::max(1, std::min(n_gpus, avail));
        if (P < 2 || !out_ms || count <= 0 || reps <= 0) return -1;
        std::vector<int> devs(P);
        for (int p = 0; p < P; ++p) devs[p] = p;
        RankGroup group(Transport::Nccl, devs);
        std::vector<double> rs(P, 0.0), ar(P, 0.0);
        group.run([&](int p, int device, Collective *coll) {
            HPR_CUDA_CHECK(cudaSetDevice(device));
            cudaStream_t st;
            HPR_CUDA_CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
            const size_t block = (((size_t)count + P - 1) / P + 63) / 64 * 64;
            double *buf = nullptr;
            HPR_CUDA_CHECK(cudaMalloc(&buf, sizeof(double) * block * P));
            HPR_CUDA_CHECK(cudaMemset(buf, 0, sizeof(double) * block * P));
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0); cudaEventCreate(&e1);
            auto timed = [&](auto &&op) {
                for (int i = 0; i < 5; ++i) op();
                HPR_CUDA_CHECK(cudaStreamSynchronize(st));
                cudaEventRecord(e0, st);
                for (int i = 0; i < reps; ++i) op();
                cudaEventRecord(e1, st);
                HPR_CUDA_CHECK(cudaEventSynchronize(e1));
                float t = 0.f;
                cudaEventElapsedTime(&t, e0, e1);
                return (double)t / reps;
            };
            rs[p] = timed([&] { coll->reduce_scatter_inplace(buf, block, st); coll->all_gather_inplace(buf, block, st); });
            ar[p] = timed([&] { coll->all_reduce(buf, (size_t)count, false, st); });
            cudaFree(buf); cudaStreamDestroy(st); cudaEventDestroy(e0); cudaEventDestroy(e1);
        });
        out_ms[0] = *std::max_element(rs.begin(), rs.end());
        out_ms[1] = *std::max_element(ar.begin(), ar.end());
        return 0;
    });
}
