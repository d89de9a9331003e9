// presolve.cpp -- bridge to the PSLP presolver (third-party, Apache-2.0, v0.0.8; the reference vendors it under
// third_party/PSLP and drives it from src/pslp_integration.cpp).  Host-side and OUT OF SCOPE for the rebuild
// (SURVEY.md 2.1 rows 12-13): PSLP is an external dependency that is compiled from where its sources lie
// (PSLP_DIR in the Makefile, default /root/reference/third_party/PSLP) and linked in when found -- its sources
// are never copied into this repository.  Without it (HPRLP_WITH_PSLP undefined) use_presolve=true solves the
// original model, which is also the reference's own fallback when presolve fails
// (src/pslp_integration.cpp:677-691).
//
// Differences from the reference bridge, by design: PSLP runs in-process (the reference forks a worker and ships
// the reduced CSR through pipes, :628-700); the calls, settings, reduced-model construction, postsolve and the
// printed original-space KKT validation (:499-624) are the same.  Running in-process is what lets solve() bring up
// the CUDA context and cuRAND on a second thread under the presolve (api.cu, warm_start).
// Crash isolation: the reference gets it from the fork -- a PSLP crash kills the worker and solve() falls back to the
// original model (:677-691).  Here the PSLP calls run inside run_guarded(): SIGSEGV / SIGBUS / SIGFPE / SIGABRT raised
// on the presolving thread are caught and turned into the same fallback (the presolver object is abandoned, never
// touched again).  Faults on PSLP's own worker threads, or memory corruption that surfaces later, are not recoverable
// this way -- a weaker guarantee than a separate process, stated here rather than hidden.
#include <chrono>
#include <cmath>
#include <csetjmp>
#include <csignal>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/hprlp_b200.h"
#include "engine.h"

#ifndef HPRLP_WITH_PSLP
namespace hpr {
bool presolve_run(const LP_info_cpu *, const HPRLP_parameters *, LP_info_cpu *, void **handle) {
    if (handle) *handle = nullptr;
    std::printf("PSLP presolve: not linked in this build; solving the original model.\n");
    return false;
}
void presolve_postsolve(HPRLP_results *, const LP_info_cpu *, void *, const HPRLP_parameters *) {}
void presolve_free(void *) {}
}  // namespace hpr
#else
#include "PSLP_API.h"
#include "PSLP_sol.h"

namespace hpr {
namespace {
struct Handle {
    Presolver *pre = nullptr;
    Settings *stg = nullptr;
    int rm = 0, rn = 0;
};

// ---- synchronous-fault guard around the third-party calls ---------------------------------------------------------
thread_local sigjmp_buf *t_guard = nullptr;   // non-null while THIS thread is inside a guarded PSLP call
void fault_handler(int sig) {
    if (t_guard) siglongjmp(*t_guard, sig);
    std::signal(sig, SIG_DFL);   // not ours (another thread, or outside a guarded region): default action
    std::raise(sig);
}
// Runs body() (plain C calls only: a non-local jump skips no destructors).  Returns 0, or the signal that aborted it.
template <class F>
int run_guarded(F &&body) {
    static std::mutex mu;   // signal dispositions are process-wide: one guarded region at a time
    std::lock_guard<std::mutex> lk(mu);
    static const int sigs[4] = {SIGSEGV, SIGBUS, SIGFPE, SIGABRT};
    struct sigaction sa, old[4];
    std::memset(&sa, 0, sizeof(sa));
    sa.sa_handler = fault_handler;
    sa.sa_flags = SA_NODEFER;
    sigemptyset(&sa.sa_mask);
    for (int i = 0; i < 4; ++i) sigaction(sigs[i], &sa, &old[i]);
    sigjmp_buf jb;
    const int caught = sigsetjmp(jb, 1);
    if (caught == 0) {
        t_guard = &jb;
        if (const char *e = std::getenv("HPRLP_TEST_PRESOLVE_FAULT")) {   // tests: a fault inside the guarded region
            if (std::strcmp(e, "abort") == 0) std::abort();
            std::raise(SIGSEGV);
        }
        body();
    }
    t_guard = nullptr;
    for (int i = 0; i < 4; ++i) sigaction(sigs[i], &old[i], nullptr);
    return caught;
}

template <typename T>
T *dup_array(const T *src, size_t count) {
    T *p = static_cast<T *>(std::malloc(sizeof(T) * (count ? count : 1)));
    if (count) std::memcpy(p, src, sizeof(T) * count);
    return p;
}
}  // namespace

bool presolve_run(const LP_info_cpu *model, const HPRLP_parameters *param, LP_info_cpu *reduced, void **handle_out) {
    if (!model || !model->A || !reduced || !handle_out) return false;
    *handle_out = nullptr;
    std::memset(reduced, 0, sizeof(*reduced));
    std::printf("Doing presolve (PSLP)...\n");
    const auto t0 = std::chrono::steady_clock::now();
    Handle *h = new Handle;
    h->stg = default_settings();
    if (!h->stg) { delete h; return false; }
    h->stg->verbose = false;   // reference src/pslp_integration.cpp:231-234
    if (param && std::isfinite(param->time_limit) && param->time_limit > 0.0)
        h->stg->max_time = std::min(h->stg->max_time, static_cast<double>(param->time_limit));
    Settings *stg = h->stg;
    Presolver *volatile pre = nullptr;   // written inside the guarded region, read after a possible siglongjmp
    const int fault = run_guarded([&] {
        pre = new_presolver(model->A->value, model->A->colIndex, model->A->rowPtr, (size_t)model->m, (size_t)model->n,
                            (size_t)model->A->numElements, model->AL, model->AU, model->l, model->u, model->c, stg);
        if (pre) run_presolver(pre);
    });
    if (fault) {   // PSLP crashed: abandon its state (never freed, never touched) and solve the original model
        std::fprintf(stderr, "[warn] PSLP presolve crashed (signal %d); solving original model\n", fault);
        delete h;
        return false;
    }
    h->pre = pre;
    auto fail = [&](const char *msg) {
        std::fprintf(stderr, "[warn] %s; solving original model\n", msg);
        if (h->pre) free_presolver(h->pre);
        free_settings(h->stg);
        delete h;
        return false;
    };
    if (!h->pre) return fail("PSLP presolver could not be created");
    std::printf("PSLP presolve time: %g seconds\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    PresolvedProblem *rp = h->pre->reduced_prob;
    if (!rp) return fail("PSLP did not return a reduced problem");
    // reduced model = build_model_from_arrays on the PSLP output (reference :341-389; m, n, nnz must be > 0)
    if (rp->m == 0 || rp->n == 0 || rp->nnz == 0 || rp->Ap[0] != 0 || rp->Ap[rp->m] != (int)rp->nnz)
        return fail("Failed to build reduced HPRLP model from PSLP output");
    reduced->m = (int)rp->m; reduced->n = (int)rp->n;
    reduced->A = static_cast<sparseMatrix *>(std::malloc(sizeof(sparseMatrix)));
    reduced->A->row = (int)rp->m; reduced->A->col = (int)rp->n; reduced->A->numElements = (int)rp->nnz;
    reduced->A->rowPtr = dup_array(rp->Ap, rp->m + 1);
    reduced->A->colIndex = dup_array(rp->Ai, rp->nnz);
    reduced->A->value = dup_array(rp->Ax, rp->nnz);
    reduced->AL = dup_array(rp->lhs, rp->m); reduced->AU = dup_array(rp->rhs, rp->m);
    reduced->l = dup_array(rp->lbs, rp->n); reduced->u = dup_array(rp->ubs, rp->n); reduced->c = dup_array(rp->c, rp->n);
    reduced->obj_constant = model->obj_constant + rp->obj_offset;
    std::printf("problem information: nRow = %d, nCol = %d, nnz A = %d\n\n", reduced->m, reduced->n, reduced->A->numElements);
    h->rm = reduced->m; h->rn = reduced->n;
    *handle_out = h;
    std::printf("PSLP presolve reduced problem: (%d, %d) -> (%d, %d)\n", model->m, model->n, reduced->m, reduced->n);
    return true;
}

namespace {
// original-space KKT validation, reference compute_original_kkt_metrics (:499-601)
void print_original_kkt(const LP_info_cpu *M, const double *x, const double *y, const double *z, const HPRLP_parameters *param) {
    const int m = M->m, n = M->n;
    std::vector<double> yp(y, y + m), zp(z, z + n), Ax(m, 0.0), ATy(n, 0.0);
    auto project = [](double &v, double lo, double hi) {
        const bool lower_inf = std::isinf(lo) && lo < 0.0, upper_inf = std::isinf(hi) && hi > 0.0;
        if (lower_inf && upper_inf) v = 0.0;
        else if (upper_inf) v = std::max(v, 0.0);
        else if (lower_inf) v = std::min(v, 0.0);
    };
    for (int i = 0; i < m; ++i) project(yp[i], M->AL[i], M->AU[i]);
    for (int j = 0; j < n; ++j) project(zp[j], M->l[j], M->u[j]);
    const sparseMatrix *A = M->A;
    for (int i = 0; i < m; ++i) {
        double s = 0.0;
        for (int k = A->rowPtr[i]; k < A->rowPtr[i + 1]; ++k) {
            s += A->value[k] * x[A->colIndex[k]];
            ATy[A->colIndex[k]] += A->value[k] * yp[i];
        }
        Ax[i] = s;
    }
    double nb2 = 0.0, nc2 = 0.0, eax = 0.0, ex = 0.0, dr = 0.0, p_lin = 0.0, d_lin = 0.0;
    for (int i = 0; i < m; ++i) {
        const double lo = std::isfinite(M->AL[i]) ? std::abs(M->AL[i]) : 0.0, hi = std::isfinite(M->AU[i]) ? std::abs(M->AU[i]) : 0.0;
        const double r = std::max(lo, hi);
        nb2 += r * r;
        double v = 0.0;
        if (std::isfinite(M->AL[i]) && Ax[i] < M->AL[i]) v = std::max(v, M->AL[i] - Ax[i]);
        if (std::isfinite(M->AU[i]) && Ax[i] > M->AU[i]) v = std::max(v, Ax[i] - M->AU[i]);
        eax += v * v;
        const double sup = yp[i] >= 0.0 ? (std::isfinite(M->AL[i]) ? M->AL[i] : 0.0) : (std::isfinite(M->AU[i]) ? M->AU[i] : 0.0);
        d_lin += yp[i] * sup;
    }
    for (int j = 0; j < n; ++j) {
        nc2 += M->c[j] * M->c[j];
        double v = 0.0;
        if (std::isfinite(M->l[j]) && x[j] < M->l[j]) v = std::max(v, M->l[j] - x[j]);
        if (std::isfinite(M->u[j]) && x[j] > M->u[j]) v = std::max(v, x[j] - M->u[j]);
        ex += v * v;
        const double r = M->c[j] - ATy[j] - zp[j];
        dr += r * r;
        p_lin += M->c[j] * x[j];
        const double sup = zp[j] >= 0.0 ? (std::isfinite(M->l[j]) ? M->l[j] : 0.0) : (std::isfinite(M->u[j]) ? M->u[j] : 0.0);
        d_lin += zp[j] * sup;
    }
    const double primal_feas = std::max(std::sqrt(eax), std::sqrt(ex)) / (1.0 + std::sqrt(nb2));
    const double dual_feas = std::sqrt(dr) / (1.0 + std::sqrt(nc2));
    const double gap = std::abs(d_lin - p_lin) / (1.0 + std::abs(d_lin) + std::abs(p_lin));
    const double tol = param ? param->stop_tol : 1e-4;
    if (std::max(primal_feas, std::max(dual_feas, gap)) <= tol) {
        std::printf("Postsolve original KKT check passed\n");
        return;
    }
    std::printf("Warning: postsolve original KKT check failed (but the primal solution and objective are reliable): ");
    bool first = true;
    if (primal_feas > tol) { std::printf("primal recover failed"); first = false; }
    if (dual_feas > tol || gap > tol) std::printf("%sdual recover failed", first ? "" : "; ");
    std::printf("\nStop Tolerance: %g\nPrimal Objective: %g\nDual Objective: %g\nPrimal Residual: %g\nDual Residual: %g\nRelative Gap: %g\n",
                tol, p_lin + M->obj_constant, d_lin + M->obj_constant, primal_feas, dual_feas, gap);
}
}  // namespace

void presolve_postsolve(HPRLP_results *result, const LP_info_cpu *original, void *handle, const HPRLP_parameters *param) {
    if (!result || !original || !handle || !result->x || !result->y || !result->z) return;
    Handle *h = static_cast<Handle *>(handle);
    std::printf("\n================================================================================\nPSLP POSTSOLVE\n"
                "================================================================================\n");
    Presolver *pre = h->pre;
    double *rx = result->x, *ry = result->y, *rz = result->z;
    const int fault = run_guarded([&] { postsolve(pre, rx, ry, rz); });
    if (fault) {   // no original-space solution can be recovered: report it (the reference's worker would have died too)
        std::fprintf(stderr, "[error] PSLP postsolve crashed (signal %d)\n", fault);
        h->pre = nullptr;   // abandoned
        std::free(result->x); std::free(result->y); std::free(result->z);
        result->x = result->y = result->z = nullptr;
        std::memset(result->status, 0, sizeof(result->status));
        std::strncpy(result->status, "ERROR", sizeof(result->status) - 1);
        return;
    }
    const Solution *sol = h->pre->sol;
    if (!sol || (int)sol->dim_x != original->n || (int)sol->dim_y != original->m) {
        std::fprintf(stderr, "[warn] PSLP postsolve returned inconsistent solution dimensions\n");
        return;
    }
    double *x = dup_array(sol->x, sol->dim_x), *y = dup_array(sol->y, sol->dim_y), *z = dup_array(sol->z, sol->dim_x);
    std::free(result->x); std::free(result->y); std::free(result->z);
    result->x = x; result->y = y; result->z = z;
    if (std::strcmp(result->status, "OPTIMAL") != 0) {
        std::printf("Skipping postsolve original KKT check since the reduced solution is not optimal\n");
        return;
    }
    print_original_kkt(original, x, y, z, param);
}

void presolve_free(void *handle) {
    if (!handle) return;
    Handle *h = static_cast<Handle *>(handle);
    if (h->pre) free_presolver(h->pre);
    if (h->stg) free_settings(h->stg);
    delete h;
}
}  // namespace hpr
#endif

// test hook: run only the presolve step and hand back the reduced model (freed with hprlp_b200_presolve_free)
extern "C" int hprlp_b200_presolve(const LP_info_cpu *model, const HPRLP_parameters *param, LP_info_cpu *reduced, void **handle) {
    return hpr::presolve_run(model, param, reduced, handle) ? 1 : 0;
}
extern "C" void hprlp_b200_presolve_free(void *handle, LP_info_cpu *reduced) {
    hpr::presolve_free(handle);
    if (reduced) hpr::free_lp_info_cpu(reduced);
}
