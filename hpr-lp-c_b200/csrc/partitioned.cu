// partitioned.cu -- one LP row-block partitioned over the GPUs of one node (new functionality: the
// reference is single-GPU).  GPU p owns a contiguous block of rows of A (A_p as CSR and A_p^T as CSR, built on the
// device), the matching blocks of y / AL / AU / row_norm, and a replica of every x-side vector.  Per iteration:
//   x-phase  w_p = A_p^T y_p (local SpMV)  ->  NCCL all-reduce of w over NVLink  ->  replicated x-update
//   y-phase  fused SpMV + projection + Halpern on the local rows (no communication)
// Residual / restart scalars that are sums over rows are all-reduced (<= 4 doubles per check); column statistics
// of the scaling and the power iteration's A^T q are all-reduced n-vectors.  One host thread + stream per GPU
// (ncclCommInitAll); every thread runs the same deterministic host logic on identical reduced scalars.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/HPRLP.h"
#include "../../include/hprlp_b200.h"
#include "engine.h"

using namespace hpr;

extern "C" HPRLP_results hprlp_b200_solve_partitioned(const LP_info_cpu *model, const HPRLP_parameters *param_in, int n_gpus,
                                                       int quiet, hprlp_b200_info *info) {
    HPRLP_parameters def;
    const HPRLP_parameters param = param_in ? *param_in : def;
    if (!model || !model->A) {
        HPRLP_results r;
        std::memset(r.status, 0, sizeof(r.status));
        std::strncpy(r.status, "ERROR", sizeof(r.status) - 1);
        r.residuals = r.primal_obj = r.gap = 0.0;
        return r;
    }
    int avail = 0;
    if (cudaGetDeviceCount(&avail) != cudaSuccess || avail < 1) throw std::runtime_error("no CUDA device");
    const int m = model->m, n = model->n;
    const int P = std::max(1, std::min(std::min(n_gpus, avail - param.device_number), m));
    if (P == 1) return hprlp_b200_solve_ex(model, &param, nullptr, 0, nullptr, nullptr, nullptr, nullptr, quiet, info);

    // row blocks balanced by nonzeros
    const int *rp = model->A->rowPtr;
    const long long nnz = model->A->numElements;
    std::vector<int> b(P + 1, 0);
    b[P] = m;
    for (int p = 1; p < P; ++p) {
        const long long target = nnz * p / P;
        b[p] = (int)(std::lower_bound(rp, rp + m + 1, (int)target) - rp);
        b[p] = std::max(b[p], b[p - 1]);
        b[p] = std::min(b[p], m);
    }
    std::vector<int> devs(P);
    for (int p = 0; p < P; ++p) devs[p] = param.device_number + p;
    std::vector<NcclComm> comms(P, nullptr);
    {
        const int rc = nccl().CommInitAll(comms.data(), P, devs.data());
        if (rc != 0) throw std::runtime_error(std::string("ncclCommInitAll failed: ") + nccl().GetErrorString(rc));
    }
    if (!quiet) {
        std::printf("Row-block partition over %d GPUs (NCCL all-reduce of A^T y per iteration): rows", P);
        for (int p = 0; p <= P; ++p) std::printf(" %d", b[p]);
        std::printf("\n");
    }

    std::vector<HPRLP_results> results(P);
    std::vector<SolveHooks> hooks(P);
    std::vector<std::string> errors(P);
    std::vector<int> lanes(2 * P, 0), nbands(P, 0);
    std::vector<std::thread> workers;
    for (int p = 0; p < P; ++p) {
        workers.emplace_back([&, p]() {
            try {
                const int mp = b[p + 1] - b[p];
                std::vector<int> rp_local((size_t)mp + 1);
                for (int i = 0; i <= mp; ++i) rp_local[i] = rp[b[p] + i] - rp[b[p]];
                sparseMatrix Ap{mp, n, rp_local[mp], model->A->colIndex + rp[b[p]], rp_local.data(), model->A->value + rp[b[p]]};
                LP_info_cpu shard{};
                shard.m = mp; shard.n = n; shard.A = &Ap;
                shard.AL = model->AL + b[p]; shard.AU = model->AU + b[p];
                shard.c = model->c; shard.l = model->l; shard.u = model->u;
                shard.obj_constant = model->obj_constant;
                HPRLP_parameters pp = param;
                pp.device_number = devs[p];
                Engine eng;
                eng.comm = comms[p]; eng.nranks = P; eng.rank = p; eng.m_global = m; eng.row0 = b[p];
                hooks[p].quiet = quiet != 0 || p != 0;
                eng.upload(&shard, devs[p]);
                eng.scale(&pp);
                results[p] = eng.solve(&pp, &hooks[p]);
                lanes[2 * p] = eng.A.G; lanes[2 * p + 1] = eng.AT.G; nbands[p] = (int)eng.A.bands.size();
            } catch (const std::exception &e) {
                errors[p] = e.what();
            }
        });
    }
    for (auto &w : workers) w.join();
    for (int p = 0; p < P; ++p) nccl().CommDestroy(comms[p]);
    for (int p = 0; p < P; ++p)
        if (!errors[p].empty()) throw std::runtime_error("partitioned solve, GPU " + std::to_string(p) + ": " + errors[p]);

    HPRLP_results out = results[0];
    double *y = static_cast<double *>(std::malloc(sizeof(double) * (size_t)m));
    for (int p = 0; p < P; ++p) {
        if (results[p].y) std::memcpy(y + b[p], results[p].y, sizeof(double) * (size_t)(b[p + 1] - b[p]));
        std::free(results[p].y);
        if (p > 0) { std::free(results[p].x); std::free(results[p].z); }
    }
    out.y = y;
    if (info) {
        const SolveHooks &h = hooks[0];
        info->lambda_max = h.lambda_max; info->sigma = h.sigma; info->setup_seconds = h.setup_seconds;
        info->scaling_seconds = h.scaling_seconds; info->power_seconds = h.power_seconds; info->loop_device_ms = h.loop_device_ms;
        info->restarts = h.restarts; info->power_iters = h.power_iters; info->kernel_launches = h.kernel_launches;
        info->b_scale = h.scal[0]; info->c_scale = h.scal[1]; info->norm_b = h.scal[2]; info->norm_c = h.scal[3];
        info->norm_b_org = h.scal[4]; info->norm_c_org = h.scal[5];
        info->lanes_A = lanes[0]; info->lanes_AT = lanes[1]; info->items_A = P; info->items_AT = P; info->bands_A = nbands[0]; info->reserved0 = 0;
    }
    return out;
}

// Diagnostic: average time (ms) of one ncclAllReduce of `count` doubles issued from one host thread per GPU,
// exactly as the partitioned solver issues it (tools/partition_study.py reports it next to the solver numbers).
extern "C" double hprlp_b200_nccl_allreduce_ms(int n_gpus, long long count, int reps) {
    int avail = 0;
    cudaGetDeviceCount(&avail);
    const int P = std::max(1, std::min(n_gpus, avail));
    if (P < 2) return 0.0;
    std::vector<int> devs(P);
    for (int p = 0; p < P; ++p) devs[p] = p;
    std::vector<NcclComm> comms(P, nullptr);
    if (nccl().CommInitAll(comms.data(), P, devs.data()) != 0) return -1.0;
    std::vector<double> ms(P, 0.0);
    std::vector<std::thread> workers;
    for (int p = 0; p < P; ++p) {
        workers.emplace_back([&, p]() {
            cudaSetDevice(devs[p]);
            cudaStream_t st;
            cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
            double *buf = nullptr;
            cudaMalloc(&buf, sizeof(double) * (size_t)count);
            cudaMemset(buf, 0, sizeof(double) * (size_t)count);
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0); cudaEventCreate(&e1);
            for (int i = 0; i < 5; ++i) nccl().AllReduce(buf, buf, (size_t)count, kNcclFloat64, kNcclSum, comms[p], st);
            cudaStreamSynchronize(st);
            cudaEventRecord(e0, st);
            for (int i = 0; i < reps; ++i) nccl().AllReduce(buf, buf, (size_t)count, kNcclFloat64, kNcclSum, comms[p], st);
            cudaEventRecord(e1, st);
            cudaEventSynchronize(e1);
            float t = 0.f;
            cudaEventElapsedTime(&t, e0, e1);
            ms[p] = t / reps;
            cudaFree(buf); cudaStreamDestroy(st); cudaEventDestroy(e0); cudaEventDestroy(e1);
        });
    }
    for (auto &w : workers) w.join();
    for (int p = 0; p < P; ++p) nccl().CommDestroy(comms[p]);
    return *std::max_element(ms.begin(), ms.end());
}
