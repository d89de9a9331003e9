// api.cu -- the C ABI of libhprlp: the seven reference entry points (include/HPRLP.h,
// include/batched_solver.h) plus the extended step-wise API (include/hprlp_b200.h), all backed by
// hpr::Engine.  Model layer = reference src/HPRLP.cu:321-537, re-written.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <atomic>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/HPRLP.h"
#include "../../include/hprlp_b200.h"
#include "../../include/version.h"
#include "abi_guard.h"
#include "engine.h"
#include <cuda_profiler_api.h>

using hpr::Engine;
using hpr::SolveHooks;

namespace {

double now_seconds() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

HPRLP_results make_error_result(const char *) { return hpr::abi_error_result(); }

void print_banner_and_params(const HPRLP_parameters *p) {   // reference src/HPRLP.cu:15-46
    std::printf("\n==================================================================\n");
    std::printf("                          HPR-LP Solver                           \n");
    std::printf("     Halpern Peaceman-Rachford Linear Programming Solver          \n");
    std::printf("                                                                  \n");
    std::printf("  Version: %s  (%s)\n", HPRLP_VERSION_STRING, HPRLP_ENGINE_STRING);
    std::printf("                                                                  \n");
    std::printf("==================================================================\n\n");
    std::printf("Solver Parameters:\n");
    std::printf("  Device:              GPU %d\n", p->device_number);
    std::printf("  Max Iterations:      %d\n", p->max_iter);
    std::printf("  Stopping Tolerance:  %.1e\n", p->stop_tol);
    std::printf("  Time Limit:          %.1f seconds\n", p->time_limit);
    std::printf("  Check Interval:      %d iterations\n", p->check_iter);
    std::printf("  cuSPARSE Only:       %s\n", p->CUSPARSE_spmv ? "Enabled (ignored: single hand-written backend)" : "Disabled");
    std::printf("  Autotune Verbose:    %s\n", p->autotune_verbose ? "Enabled (ignored)" : "Disabled");
    std::printf("  PSLP Presolve:       %s\n", p->use_presolve ? "Enabled" : "Disabled");
    std::printf("  Scaling:\n");
    std::printf("    - Curtis-Reid:     %s\n", p->use_CR_scaling ? "Enabled" : "Disabled");
    std::printf("    - Ruiz:            %s\n", p->use_Ruiz_scaling ? "Enabled" : "Disabled");
    std::printf("    - Pock-Chambolle:  %s\n", p->use_Pock_Chambolle_scaling ? "Enabled" : "Disabled");
    std::printf("    - Bounds/Cost:     %s\n\n", p->use_bc_scaling ? "Enabled" : "Disabled");
}

// setup + scaling of HPRLP_main_solve (reference src/HPRLP.cu:124-147)
void prepare_engine(Engine &eng, const LP_info_cpu *lp, const HPRLP_parameters *param, SolveHooks *hooks, bool quiet) {
    const double t0 = now_seconds();
    eng.upload(lp, param->device_number);
    hooks->setup_seconds = now_seconds() - t0;
    if (!quiet) std::printf("Setup (copy and allocation) time = %.2f seconds\n", hooks->setup_seconds);
    const double t1 = now_seconds();
    eng.scale(param);
    hooks->scaling_seconds = now_seconds() - t1;
    if (!quiet) std::printf("Scaling time = %.2f seconds\n", hooks->scaling_seconds);
}

void fill_info(const Engine &eng, const SolveHooks &h, hprlp_b200_info *info) { hpr::fill_b200_info(eng, h, info); }

}  // namespace

// ---- first-solve warm-up: one background thread per device, started at most once per process ------------------------
namespace {
std::mutex g_warm_mu;
std::thread g_warm_thread[64];
bool g_warm_started[64] = {};
void warm_start(int device) {
    if (device < 0 || device >= 64 || getenv("HPRLP_NO_WARMUP")) return;
    std::lock_guard<std::mutex> lk(g_warm_mu);
    if (g_warm_started[device]) return;
    g_warm_started[device] = true;
    g_warm_thread[device] = std::thread([device] { hpr::warm_device(device); });
}
void warm_wait(int device) {
    if (device < 0 || device >= 64) return;
    std::thread t;
    {
        std::lock_guard<std::mutex> lk(g_warm_mu);
        if (g_warm_thread[device].joinable()) t = std::move(g_warm_thread[device]);
    }
    if (t.joinable()) t.join();
}
struct WarmJoiner {   // no joinable std::thread may be left at process exit
    ~WarmJoiner() { for (int d = 0; d < 64; ++d) warm_wait(d); }
} g_warm_joiner;
}  // namespace

struct hprlp_b200_engine {
    Engine eng;
    HPRLP_parameters param;
    SolveHooks hooks;
    bool finished = false;
};

extern "C" {

HPRLP_results HPRLP_main_solve(const LP_info_cpu *lp, const HPRLP_parameters *param) {
    return hprlp_b200_solve_ex(lp, param, nullptr, 0, nullptr, nullptr, nullptr, nullptr, 0, nullptr);
}

static HPRLP_results solve_ex_impl(const LP_info_cpu *lp, const HPRLP_parameters *param_in, const double *power_z0,
                                   int n_trace, const int *trace_iters, double *trace_x, double *trace_y,
                                   double *trace_z, int quiet, hprlp_b200_info *info);

HPRLP_results hprlp_b200_solve_ex(const LP_info_cpu *lp, const HPRLP_parameters *param_in, const double *power_z0,
                                  int n_trace, const int *trace_iters, double *trace_x, double *trace_y,
                                  double *trace_z, int quiet, hprlp_b200_info *info) {
    // CUDA failures (out of memory, ...) become an "ERROR" result: nothing is thrown across the C ABI
    return hpr::abi_guard_results("HPRLP_main_solve", [&] {
        return solve_ex_impl(lp, param_in, power_z0, n_trace, trace_iters, trace_x, trace_y, trace_z, quiet, info);
    });
}

static HPRLP_results solve_ex_impl(const LP_info_cpu *lp, const HPRLP_parameters *param_in, const double *power_z0,
                                   int n_trace, const int *trace_iters, double *trace_x, double *trace_y,
                                   double *trace_z, int quiet, hprlp_b200_info *info) {
    if (!lp || !lp->A) {
        std::cerr << "[error] Null model pointer" << std::endl;
        return make_error_result("ERROR");
    }
    HPRLP_parameters def;
    const HPRLP_parameters *param = param_in ? param_in : &def;
    warm_wait(param->device_number);   // a warm-up started earlier (hprlp_b200_warmup / solve with presolve) must be through
    if (!quiet) print_banner_and_params(param);
    static const bool timing = getenv("HPRLP_TIMING") != nullptr;   // stage wall times (with device syncs) on stderr
    const double t0 = now_seconds();
    double t1 = t0, t2 = t0;
    HPRLP_results out;
    {
        Engine eng;
        SolveHooks hooks;
        hooks.power_z0 = power_z0;
        hooks.n_trace = n_trace; hooks.trace_iters = trace_iters;
        hooks.trace_x = trace_x; hooks.trace_y = trace_y; hooks.trace_z = trace_z;
        hooks.quiet = quiet != 0;
        prepare_engine(eng, lp, param, &hooks, hooks.quiet);
        if (timing) { cudaStreamSynchronize(eng.stream); t1 = now_seconds(); }
        out = eng.solve(param, &hooks);
        fill_info(eng, hooks, info);
        t2 = now_seconds();
    }
    if (timing)
        std::fprintf(stderr, "[hprlp timing] upload+scale %.4f s, solve (power + loop + collect) %.4f s, teardown %.4f s\n", t1 - t0, t2 - t1,
                     now_seconds() - t2);
    return out;
}

int hprlp_b200_power_start(int m, int device, double *out) {
    if (m <= 0 || !out) return -1;
    return hpr::abi_guard_int("hprlp_b200_power_start", [&]() -> int {
    LP_info_cpu lp{};
    // minimal 1-nnz model just to own a stream/buffers of length m
    std::vector<int> rp((size_t)m + 1, 1); rp[0] = 0;
    int ci = 0; double v = 1.0;
    sparseMatrix A{m, 1, 1, &ci, rp.data(), &v};
    std::vector<double> zm((size_t)m, 0.0); double z1 = 0.0;
    lp.m = m; lp.n = 1; lp.A = &A; lp.AL = zm.data(); lp.AU = zm.data(); lp.c = &z1; lp.l = &z1; lp.u = &z1;
    Engine eng;
    eng.upload(&lp, device);
    eng.power_start_vector(eng.wm);
    HPR_CUDA_CHECK(cudaMemcpyAsync(out, eng.wm, sizeof(double) * m, cudaMemcpyDeviceToHost, eng.stream));
    HPR_CUDA_CHECK(cudaStreamSynchronize(eng.stream));
    return 0;
    });
}

hprlp_b200_engine *hprlp_b200_engine_create(const LP_info_cpu *lp, const HPRLP_parameters *param_in) {
    if (!lp || !lp->A) return nullptr;
    return hpr::abi_guard<hprlp_b200_engine *>("hprlp_b200_engine_create", [&]() -> hprlp_b200_engine * {
        HPRLP_parameters def;
        std::unique_ptr<hprlp_b200_engine> h(new hprlp_b200_engine);   // released (engine, device memory) if setup throws
        h->param = param_in ? *param_in : def;
        h->hooks.quiet = true;
        prepare_engine(h->eng, lp, &h->param, &h->hooks, true);
        h->eng.solve_begin(&h->param, &h->hooks);
        return h.release();
    }, [] { return (hprlp_b200_engine *)nullptr; });
}

// Resident engine of ONE RANK of a row-partitioned solve (one process per GPU): this process uploads row block `rank`
// of the model to device param->device_number; all ranks then call hprlp_b200_engine_run with the same arguments.
hprlp_b200_engine *hprlp_b200_engine_create_rank(const LP_info_cpu *lp, const HPRLP_parameters *param_in, hprlp_b200_comm *comm) {
    if (!lp || !lp->A) return nullptr;
    if (!comm || !comm->coll) return hprlp_b200_engine_create(lp, param_in);
    if (comm->coll->nranks > lp->m) return nullptr;
    return hpr::abi_guard<hprlp_b200_engine *>("hprlp_b200_engine_create_rank", [&]() -> hprlp_b200_engine * {
        HPRLP_parameters def;
        std::unique_ptr<hprlp_b200_engine> h(new hprlp_b200_engine);
        h->param = param_in ? *param_in : def;
        h->param.device_number = comm->device;
        h->hooks.quiet = true;
        hpr::Collective *coll = comm->coll.get();
        const std::vector<int> b = hpr::row_blocks_by_nnz(lp->A->rowPtr, lp->m, coll->nranks);
        h->eng.set_partition(coll, lp->m, b[coll->rank]);
        const double t0 = now_seconds();
        hpr::upload_row_block(h->eng, lp, b[coll->rank], b[coll->rank + 1], comm->device);
        h->hooks.setup_seconds = now_seconds() - t0;
        const double t1 = now_seconds();
        h->eng.scale(&h->param);
        h->hooks.scaling_seconds = now_seconds() - t1;
        h->eng.solve_begin(&h->param, &h->hooks);
        return h.release();
    }, [] { return (hprlp_b200_engine *)nullptr; });
}

double hprlp_b200_engine_run(hprlp_b200_engine *h, int iters) {
    if (!h || h->finished || iters <= 0) return 0.0;
    return hpr::abi_guard_double("hprlp_b200_engine_run", [&]() -> double {
    Engine &e = h->eng;
    cudaEvent_t e0, e1;
    HPR_CUDA_CHECK(cudaEventCreate(&e0));
    HPR_CUDA_CHECK(cudaEventCreate(&e1));
    HPR_CUDA_CHECK(cudaEventRecord(e0, e.stream));
    h->finished = e.solve_advance(&h->param, &h->hooks, e.loop.iter + iters);
    HPR_CUDA_CHECK(cudaEventRecord(e1, e.stream));
    HPR_CUDA_CHECK(cudaEventSynchronize(e1));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (h->finished) {   // the bench never wants the solution vectors
        std::free(e.loop.output.x); std::free(e.loop.output.y); std::free(e.loop.output.z);
        e.loop.output.x = e.loop.output.y = e.loop.output.z = nullptr;
    }
    return (double)ms;
    });
}

double hprlp_b200_engine_time_phase(hprlp_b200_engine *h, int which, int reps) {
    if (!h || reps <= 0) return 0.0;
    return hpr::abi_guard_double("hprlp_b200_engine_time_phase", [&]() -> double { return h->eng.time_phase_ms(which, reps); });
}

int hprlp_b200_engine_residuals(hprlp_b200_engine *h, double *kkt, double *pobj, double *dobj) {
    if (!h) return -1;
    const hpr::Residuals &r = h->eng.loop.res;
    if (kkt) *kkt = r.kkt;
    if (pobj) *pobj = r.primal_obj;
    if (dobj) *dobj = r.dual_obj;
    return h->eng.loop.iter;
}

void hprlp_b200_engine_info(hprlp_b200_engine *h, hprlp_b200_info *info) {
    if (!h || !info) return;
    h->eng.fill_hooks(&h->hooks);
    fill_info(h->eng, h->hooks, info);
}

void hprlp_b200_engine_destroy(hprlp_b200_engine *h) { delete h; }

// Device memory cached by finished solves (the engines' private stream-ordered pool) goes back to the driver.
void hprlp_b200_release_cached_memory(void) {
    hpr::abi_guard_int("hprlp_b200_release_cached_memory", [&]() -> int { hpr::release_cached_device_memory(); return 0; });
}

int hprlp_b200_scale_only(const LP_info_cpu *lp, const HPRLP_parameters *param_in, double *A_val, int *AT_rowPtr,
                          int *AT_col, double *AT_val, double *AL, double *AU, double *l, double *u, double *c,
                          double *row_norm, double *col_norm, double *scalars6) {
    if (!lp || !lp->A) return -1;
    return hpr::abi_guard_int("hprlp_b200_scale_only", [&]() -> int {
    HPRLP_parameters def;
    const HPRLP_parameters *param = param_in ? param_in : &def;
    Engine eng;
    eng.upload(lp, param->device_number);
    eng.scale(param);
    const int m = eng.m, n = eng.n;
    const size_t nnz = (size_t)eng.nnz;
    auto d2h = [&](void *dst, const void *src, size_t bytes) {
        if (dst) HPR_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, eng.stream));
    };
    d2h(A_val, eng.A.val, sizeof(double) * nnz);
    d2h(AT_rowPtr, eng.AT.rowPtr, sizeof(int) * ((size_t)n + 1));
    d2h(AT_col, eng.AT.col, sizeof(int) * nnz);
    d2h(AT_val, eng.AT.val, sizeof(double) * nnz);
    d2h(AL, eng.AL, sizeof(double) * m); d2h(AU, eng.AU, sizeof(double) * m);
    d2h(l, eng.l, sizeof(double) * n); d2h(u, eng.u, sizeof(double) * n); d2h(c, eng.c, sizeof(double) * n);
    d2h(row_norm, eng.row_norm, sizeof(double) * m); d2h(col_norm, eng.col_norm, sizeof(double) * n);
    HPR_CUDA_CHECK(cudaStreamSynchronize(eng.stream));
    if (scalars6) {
        scalars6[0] = eng.b_scale; scalars6[1] = eng.c_scale; scalars6[2] = eng.norm_b; scalars6[3] = eng.norm_c;
        scalars6[4] = eng.norm_b_org; scalars6[5] = eng.norm_c_org;
    }
    return 0;
    });
}

// Starts the first-solve warm-up (CUDA context, module load, cuRAND) of `device` on a background thread and returns at
// once; the next solve on that device waits for it.  build/solve_mps_file calls it before parsing the MPS file.
void hprlp_b200_warmup(int device) { warm_start(device); }

void hprlp_b200_profiler_start(void) { cudaProfilerStart(); }
void hprlp_b200_profiler_stop(void) { cudaProfilerStop(); }

const char *hprlp_b200_version(void) { return "hprlp-b200 " HPRLP_ENGINE_STRING; }

// ---------------------------------------------------------------------------------------------
// model layer (reference src/HPRLP.cu:321-537)
// ---------------------------------------------------------------------------------------------
LP_info_cpu *create_model_from_arrays(int m, int n, int nnz, const int *rowPtr, const int *colIndex,
                                      const HPRLP_FLOAT *values, const HPRLP_FLOAT *AL, const HPRLP_FLOAT *AU,
                                      const HPRLP_FLOAT *l, const HPRLP_FLOAT *u, const HPRLP_FLOAT *c, bool is_csc) {
    if (m <= 0 || n <= 0 || nnz <= 0) {
        std::cerr << "[error] Invalid dimensions: m=" << m << ", n=" << n << ", nnz=" << nnz << std::endl;
        return nullptr;
    }
    if (!rowPtr || !colIndex || !values || !AL || !AU || !l || !u || !c) {
        std::cerr << "[error] Null pointer in input arrays" << std::endl;
        return nullptr;
    }
    // CSR validation as in the reference's build_model_from_arrays (src/mps_reader.cpp:1413-1421)
    const int nptr = is_csc ? n : m;
    if (rowPtr[0] != 0 || rowPtr[nptr] != nnz) {
        std::cerr << "Error: Invalid CSR format: row_ptr[0] = " << rowPtr[0] << ", row_ptr[" << nptr << "] = " << rowPtr[nptr]
                  << ", expected 0 and " << nnz << "\n";
        std::cerr << "[error] Model creation failed" << std::endl;
        return nullptr;
    }
    LP_info_cpu *model = new LP_info_cpu;
    model->m = m; model->n = n; model->obj_constant = 0.0;
    model->A = static_cast<sparseMatrix *>(std::malloc(sizeof(sparseMatrix)));
    model->A->row = m; model->A->col = n; model->A->numElements = nnz;
    model->A->rowPtr = static_cast<int *>(std::malloc(sizeof(int) * ((size_t)m + 1)));
    model->A->colIndex = static_cast<int *>(std::malloc(sizeof(int) * (size_t)nnz));
    model->A->value = static_cast<double *>(std::malloc(sizeof(double) * (size_t)nnz));
    if (is_csc) {
        // CSC(A) is CSR(A^T): transposing it (stable counting sort) yields CSR(A) (src/HPRLP.cu:354-396)
        hpr::csr_transpose_host_mt(n, m, nnz, rowPtr, colIndex, values, model->A->rowPtr, model->A->colIndex, model->A->value);
    } else {
        std::memcpy(model->A->rowPtr, rowPtr, sizeof(int) * ((size_t)m + 1));
        hpr::copy_mt(model->A->colIndex, colIndex, sizeof(int) * (size_t)nnz);
        hpr::copy_mt(model->A->value, values, sizeof(double) * (size_t)nnz);
    }
    std::printf("problem information: nRow = %d, nCol = %d, nnz A = %d\n\n", m, n, nnz);
    auto dup = [](const double *src, int len) {
        double *d = static_cast<double *>(std::malloc(sizeof(double) * (size_t)len));
        std::memcpy(d, src, sizeof(double) * (size_t)len);
        return d;
    };
    model->AL = dup(AL, m); model->AU = dup(AU, m);
    model->c = dup(c, n); model->l = dup(l, n); model->u = dup(u, n);
    return model;
}

LP_info_cpu *create_model_from_mps(const char *mps_file_path) {
    if (!mps_file_path) {
        std::cerr << "[error] Null MPS file path pointer" << std::endl;
        return nullptr;
    }
    LP_info_cpu *model = new LP_info_cpu;
    model->A = nullptr; model->m = 0; model->n = 0;
    model->AL = model->AU = model->c = model->l = model->u = nullptr;
    model->obj_constant = 0.0;
    bool ok = false;
    try {
        ok = hpr::build_model_from_mps(mps_file_path, model);
    } catch (const std::exception &e) {
        std::cerr << "[error] Failed to read MPS file: " << e.what() << std::endl;
        delete model;
        return nullptr;
    }
    if (!ok || !model->A || model->m <= 0 || model->n <= 0) {
        std::cerr << "[error] Invalid model from MPS file" << std::endl;
        hpr::free_lp_info_cpu(model);
        delete model;
        return nullptr;
    }
    return model;
}

HPRLP_results solve(const LP_info_cpu *model, const HPRLP_parameters *param) {
    if (!model) {
        std::cerr << "[error] Null model pointer" << std::endl;
        return make_error_result("ERROR");
    }
    HPRLP_parameters default_param;
    const HPRLP_parameters *actual = param ? param : &default_param;
    if (!actual->use_presolve) return HPRLP_main_solve(model, actual);
    return hpr::abi_guard_results("solve", [&]() -> HPRLP_results {
    // PSLP runs in this process (no fork), so the first solve of a process can create its CUDA context and load cuRAND
    // on a second host thread WHILE the presolve runs on this one (the reference's forked worker cannot be overlapped
    // with CUDA initialisation in the parent without risking a fork of a CUDA process).  No-op from the second solve on.
    warm_start(actual->device_number);
    LP_info_cpu reduced{};
    void *handle = nullptr;
    const bool presolve_ok = hpr::presolve_run(model, actual, &reduced, &handle);
    const LP_info_cpu *solve_model = presolve_ok ? &reduced : model;
    HPRLP_results result = HPRLP_main_solve(solve_model, actual);
    if (presolve_ok) {
        if (result.x && result.y && result.z) hpr::presolve_postsolve(&result, model, handle, actual);
        hpr::presolve_free(handle);
        hpr::free_lp_info_cpu(&reduced);
    }
    return result;
    });
}

void free_model(LP_info_cpu *model) {
    if (!model) return;
    hpr::free_lp_info_cpu(model);
    delete model;
}

}  // extern "C"

namespace hpr {
void free_lp_info_cpu(LP_info_cpu *lp) {
    if (!lp) return;
    if (lp->A) {
        std::free(lp->A->rowPtr); std::free(lp->A->colIndex); std::free(lp->A->value);
        std::free(lp->A);
        lp->A = nullptr;
    }
    std::free(lp->AL); std::free(lp->AU); std::free(lp->c); std::free(lp->l); std::free(lp->u);
    lp->AL = lp->AU = lp->c = lp->l = lp->u = nullptr;
}
}  // namespace hpr
