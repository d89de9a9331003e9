/*
 * hpr_oracle.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement (plain C, fp64) of the HPR-LP
 * hot path of PolyU-IOR/HPR-LP-C.  Nothing in the product library links, calls or falls back
 * to this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg use it.
 *
 * PARITY PIN: the reference ships no tests, golden vectors or CPU path, so the
 * pins are (1) the toy LP known answer of the reference's own examples
 * (examples/cpp/example_direct_lp.cpp:14: x=(2.8,3.6), obj=-26.4), (2) iterates/objectives of the
 * reference's own CUDA build (oracle/_ref, built by oracle/build_ref.sh) recorded on a B200
 * into tests/golden/ref_*.json by tests/golden/make_ref_golden.py, and (3) live ref-vs-new
 * runs in the `-m gpu` tests.
 *
 * Every function cites the reference file:line it restates.  Sums are sequential in CSR order
 * (the reference's are cuSPARSE/shuffle trees), so agreement is to rounding, not bitwise.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int max_iter;
    double stop_tol;
    double time_limit;
    int check_iter;
    int use_cr, use_ruiz, use_pc, use_bc;
} oracle_params;

typedef struct {
    double residuals, primal_obj, dual_obj, gap, err_rp, err_rd;
    int iter;
    char status[64];
    double lambda_max, sigma;
    int restarts;
    double b_scale, c_scale, norm_b, norm_c, norm_b_org, norm_c_org;
    int power_iters;
    double solve_seconds; /* loop only, excludes setup/scaling/power */
} oracle_info;

typedef struct {
    int rows, cols, nnz;
    int *rp, *ci;
    double *v;
} csr_t;

/* ---- reference src/utils.cu:203-232 (CSR_transpose_host): stable counting sort by column ---- */
void oracle_transpose(int rows, int cols, int nnz, const int *rp, const int *ci, const double *v,
                      int *trp, int *tci, double *tv) {
    int *next = (int *)calloc((size_t)cols + 2, sizeof(int));
    for (int k = 0; k < nnz; ++k) next[ci[k] + 2]++;
    for (int j = 2; j < cols + 2; ++j) next[j] += next[j - 1];
    for (int i = 0; i < rows; ++i)
        for (int k = rp[i]; k < rp[i + 1]; ++k) {
            int pos = next[ci[k] + 1]++;
            tv[pos] = v[k];
            tci[pos] = i;
        }
    for (int j = 0; j <= cols; ++j) trp[j] = next[j];
    free(next);
}

static double nrm2(const double *x, int n) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += x[i] * x[i];
    return sqrt(s);
}
static double dot(const double *x, const double *y, int n) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += x[i] * y[i];
    return s;
}
static void spmv(const csr_t *M, const double *x, double *out) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < M->rows; ++i) {
        double s = 0.0;
        for (int k = M->rp[i]; k < M->rp[i + 1]; ++k) s = fma(M->v[k], x[M->ci[k]], s);
        out[i] = s;
    }
}

/* reference src/cuda_kernels/HPR_cuda_kernels.cu:34-43 (conceptual_b_kernel) + l2 norm */
static double norm_b_of(const double *AL, const double *AU, int m) {
    double s = 0.0;
    for (int i = 0; i < m; ++i) {
        double a = isinf(AL[i]) ? 0.0 : AL[i];
        double b = isinf(AU[i]) ? 0.0 : AU[i];
        double t = fmax(fabs(a), fabs(b));
        s += t * t;
    }
    return sqrt(s);
}

/* reference src/cuda_kernels/HPR_cuda_kernels.cu:91-120 (CSR_A_row_norm_kernel) */
static void row_norm(const csr_t *M, double *out, int kind) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < M->rows; ++i) {
        double r = 0.0;
        for (int k = M->rp[i]; k < M->rp[i + 1]; ++k) {
            double a = fabs(M->v[k]);
            if (kind == 99) { if (r < a) r = a; } else r += a;
        }
        r = sqrt(r);
        if (r < 1e-15) r = 1.0;
        out[i] = r;
    }
}

/* reference src/cuda_kernels/HPR_cuda_kernels.cu:122-157: value[k] (*|/)= rowfac[row]; then
 * value[k] (*|/)= colfac[col] -- two separately rounded operations, row factor first. */
static void scale_matrix(csr_t *M, const double *rowfac, const double *colfac, int divide) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < M->rows; ++i)
        for (int k = M->rp[i]; k < M->rp[i + 1]; ++k) {
            double t = M->v[k];
            if (divide) { t /= rowfac[i]; t /= colfac[M->ci[k]]; }
            else        { t *= rowfac[i]; t *= colfac[M->ci[k]]; }
            M->v[k] = t;
        }
}

/* reference src/scaling.cu:5-31 (curtis_reid_log_update_kernel) */
static void cr_sweep(const csr_t *M, const double *other, double *out) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < M->rows; ++i) {
        int cnt = M->rp[i + 1] - M->rp[i];
        if (cnt <= 0) { out[i] = 0.0; continue; }
        double s = 0.0;
        for (int k = M->rp[i]; k < M->rp[i + 1]; ++k)
            s += -log(fmax(fabs(M->v[k]), 1e-300)) - other[M->ci[k]];
        out[i] = s / (double)cnt;
    }
}

typedef struct {
    double *row_norm, *col_norm;
    double b_scale, c_scale, norm_b, norm_c, norm_b_org, norm_c_org;
} scaling_t;

/* reference src/scaling.cu:88-216 (scaling) and :40-84 (apply_curtis_reid_scaling).
 * A and AT are scaled with the same operation order so both copies stay bit-identical. */
static void do_scaling(csr_t *A, csr_t *AT, double *AL, double *AU, double *c, double *l, double *u,
                       const oracle_params *p, scaling_t *sc) {
    int m = A->rows, n = A->cols;
    double *t1 = (double *)calloc(m, sizeof(double));
    double *t2 = (double *)calloc(n, sizeof(double));
    for (int i = 0; i < m; ++i) sc->row_norm[i] = 1.0;
    for (int j = 0; j < n; ++j) sc->col_norm[j] = 1.0;
    sc->norm_b_org = 1.0 + norm_b_of(AL, AU, m);
    sc->norm_c_org = 1.0 + nrm2(c, n);

    if (p->use_cr) {
        for (int it = 0; it < 20; ++it) { cr_sweep(A, t2, t1); cr_sweep(AT, t1, t2); }
        for (int i = 0; i < m; ++i) t1[i] = fmin(fmax(exp(t1[i]), 1e-30), 1e30);
        for (int j = 0; j < n; ++j) t2[j] = fmin(fmax(exp(t2[j]), 1e-30), 1e30);
        for (int i = 0; i < m; ++i) sc->row_norm[i] /= t1[i];
        for (int j = 0; j < n; ++j) sc->col_norm[j] /= t2[j];
        scale_matrix(A, t1, t2, 0);
        /* AT: value *= r[col] then *= gamma[row] (scaling.cu:72-76) -- same order of factors */
#pragma omp parallel for schedule(static)
        for (int j = 0; j < n; ++j)
            for (int k = AT->rp[j]; k < AT->rp[j + 1]; ++k) {
                double t = AT->v[k]; t *= t1[AT->ci[k]]; t *= t2[j]; AT->v[k] = t;
            }
        for (int i = 0; i < m; ++i) { AL[i] *= t1[i]; AU[i] *= t1[i]; }
        for (int j = 0; j < n; ++j) { c[j] *= t2[j]; l[j] /= t2[j]; u[j] /= t2[j]; }
    }
    int rounds = (p->use_ruiz ? 10 : 0) + (p->use_pc ? 1 : 0);
    for (int it = 0; it < rounds; ++it) {
        int kind = (it < (p->use_ruiz ? 10 : 0)) ? 99 : 1;
        row_norm(A, t1, kind);   /* both norms from the matrix BEFORE either is applied (scaling.cu:127-144) */
        row_norm(AT, t2, kind);
        for (int i = 0; i < m; ++i) { sc->row_norm[i] *= t1[i]; AL[i] /= t1[i]; AU[i] /= t1[i]; }
        for (int j = 0; j < n; ++j) sc->col_norm[j] *= t2[j];
        scale_matrix(A, t1, t2, 1);
#pragma omp parallel for schedule(static)
        for (int j = 0; j < n; ++j)
            for (int k = AT->rp[j]; k < AT->rp[j + 1]; ++k) {
                double t = AT->v[k]; t /= t1[AT->ci[k]]; t /= t2[j]; AT->v[k] = t;
            }
        for (int j = 0; j < n; ++j) { c[j] /= t2[j]; l[j] *= t2[j]; u[j] *= t2[j]; }
    }
    if (p->use_bc) {
        sc->b_scale = 1.0 + norm_b_of(AL, AU, m);
        sc->c_scale = 1.0 + nrm2(c, n);
        double bs = 1.0 / sc->b_scale, cs = 1.0 / sc->c_scale;   /* cublasDscal by the reciprocal */
        for (int i = 0; i < m; ++i) { AU[i] *= bs; AL[i] *= bs; }
        for (int j = 0; j < n; ++j) { l[j] *= bs; u[j] *= bs; c[j] *= cs; }
    } else {
        sc->b_scale = 1.0; sc->c_scale = 1.0;
    }
    sc->norm_b = norm_b_of(AL, AU, m);
    sc->norm_c = nrm2(c, n);
    free(t1); free(t2);
}

/* reference src/power_iteration.cu:20-119 (power_method_cusparse).  z0 = N(0,1)+1e-8 from cuRAND in the
 * reference; here z0 is an input (the GPU tests pass the engine's own start vector). Returns lambda
 * (caller multiplies by 1.01, src/HPRLP.cu:86). */
static double power_method(const csr_t *A, const csr_t *AT, const double *z0, int max_iter, double tol,
                           int *iters_out) {
    int m = A->rows, n = A->cols;
    double *z = (double *)malloc(sizeof(double) * m), *q = (double *)malloc(sizeof(double) * m);
    double *atq = (double *)malloc(sizeof(double) * n);
    memcpy(z, z0, sizeof(double) * m);
    double lambda = 1.0;
    int it;
    for (it = 1; it <= max_iter; ++it) {
        double invn = 1.0 / sqrt(dot(z, z, m) + 2.220446049250313e-16);
        for (int i = 0; i < m; ++i) q[i] = invn * z[i];
        spmv(AT, q, atq);
        spmv(A, atq, z);
        if (it % 10 == 0) {
            lambda = dot(q, z, m);
            for (int i = 0; i < m; ++i) q[i] = -lambda * q[i] + 1.0 * z[i];
            if (nrm2(q, m) < tol) break;
        }
    }
    if (iters_out) *iters_out = it > max_iter ? max_iter : it;
    free(z); free(q); free(atq);
    return lambda;
}

/* reference src/utils.cu:100-102 (step) */
static int step_of(int iter) {
    int s = (int)(pow(10, floor(log10((double)iter))) / 10);
    return s > 10 ? s : 10;
}

typedef struct {
    int restart_flag, first_restart;
    double last_gap, current_gap, save_gap, best_gap, best_sigma;
    int inner, times;
} restart_t;

/* ------------------------------------------------------------------------------------------
 * oracle_solve: reference src/HPRLP.cu:116-311 (HPRLP_main_solve) driving
 *   src/cuda_kernels/HPR_cuda_kernels.cu:203-295 (update formulas), src/main_iterate.cu:229-420
 *   (residuals, restart, sigma, stopping), :486-515 (weighted norm), src/utils.cu:143-200 (unscale).
 * trace: after every check iteration with (iter+1) == trace_iters[t], the unscaled (x_bar,y_bar,z_bar)
 * are stored -- identical to what the reference returns for max_iter = trace_iters[t] (a multiple of 10).
 * ------------------------------------------------------------------------------------------ */
int oracle_solve(int m, int n, const int *rowPtr, const int *colIndex, const double *values,
                 const double *AL_in, const double *AU_in, const double *l_in, const double *u_in,
                 const double *c_in, double obj_constant, const oracle_params *p, const double *power_z0,
                 double *x_out, double *y_out, double *z_out, oracle_info *info,
                 int n_trace, const int *trace_iters, double *trace_x, double *trace_y, double *trace_z) {
    int nnz = rowPtr[m];
    csr_t A = {m, n, nnz, (int *)malloc(sizeof(int) * (m + 1)), (int *)malloc(sizeof(int) * nnz),
               (double *)malloc(sizeof(double) * nnz)};
    csr_t AT = {n, m, nnz, (int *)malloc(sizeof(int) * (n + 1)), (int *)malloc(sizeof(int) * nnz),
                (double *)malloc(sizeof(double) * nnz)};
    memcpy(A.rp, rowPtr, sizeof(int) * (m + 1));
    memcpy(A.ci, colIndex, sizeof(int) * nnz);
    memcpy(A.v, values, sizeof(double) * nnz);
    oracle_transpose(m, n, nnz, A.rp, A.ci, A.v, AT.rp, AT.ci, AT.v);
#define DUP(name, src, len) double *name = (double *)malloc(sizeof(double) * (len)); memcpy(name, src, sizeof(double) * (len))
    DUP(AL, AL_in, m); DUP(AU, AU_in, m); DUP(l, l_in, n); DUP(u, u_in, n); DUP(c, c_in, n);
    scaling_t sc; sc.row_norm = (double *)malloc(sizeof(double) * m); sc.col_norm = (double *)malloc(sizeof(double) * n);
    do_scaling(&A, &AT, AL, AU, c, l, u, p, &sc);

    double *z0 = (double *)malloc(sizeof(double) * m);
    for (int i = 0; i < m; ++i) z0[i] = power_z0 ? power_z0[i] : 1e-8;   /* odd-m cuRAND quirk: 0 + 1e-8 */
    int piters = 0;
    double lambda_max = power_method(&A, &AT, z0, 5000, 1e-4, &piters) * 1.01;
    free(z0);

    double sigma = (sc.norm_b > 1e-8 && sc.norm_c > 1e-8) ? sc.norm_b / sc.norm_c : 1.0;
    restart_t R; memset(&R, 0, sizeof(R));
    R.first_restart = 1; R.best_sigma = sigma;
    R.last_gap = R.current_gap = R.save_gap = R.best_gap = INFINITY;

#define VEC(name, len) double *name = (double *)calloc((len), sizeof(double))
    VEC(x, n); VEC(x0, n); VEC(x_hat, n); VEC(x_bar, n); VEC(z_bar, n); VEC(x_tmp, n);
    VEC(y, m); VEC(y0, m); VEC(y_bar, m); VEC(y_obj, m); VEC(y_tmp, m);
    VEC(w, n); VEC(ax, m);
    int k_inner = 0; /* device-side halpern_inner */
    double err_rp = 0, err_rd = 0, pobj = 0, dobj = 0, gap = 0, kkt = INFINITY;
    const double obj_scale = sc.b_scale * sc.c_scale;
    struct timespec t0; clock_gettime(CLOCK_MONOTONIC, &t0);
    const char *status = "CONTINUE";
    int iter;
    for (iter = 0;; ++iter) {
        int periodic = (iter % p->check_iter == 0);
        int compute_gap = periodic && iter > 0;
        struct timespec tn; clock_gettime(CLOCK_MONOTONIC, &tn);
        double elapsed = (tn.tv_sec - t0.tv_sec) + 1e-9 * (tn.tv_nsec - t0.tv_nsec);
        int print_flag = (iter % step_of(iter) == 0) || iter == p->max_iter || elapsed > p->time_limit;
        if (periodic || print_flag) {
            /* compute_residuals, src/main_iterate.cu:229-309 */
            pobj = obj_scale * dot(c, x_bar, n) + obj_constant;
            dobj = obj_scale * (dot(y_obj, y_bar, m) + dot(x_bar, z_bar, n)) + obj_constant;
            gap = fabs(pobj - dobj) / (1.0 + fabs(pobj) + fabs(dobj));
            double g_dot = 0, g_dy = 0, g_dx = 0;
            if (compute_gap) {
                spmv(&A, x_tmp, ax);
                g_dot = dot(ax, y_tmp, m); g_dy = dot(y_tmp, y_tmp, m); g_dx = dot(x_tmp, x_tmp, n);
            }
            spmv(&AT, y_bar, w);
            double s = 0.0;
            for (int j = 0; j < n; ++j) { double rd = (c[j] - w[j] - z_bar[j]) * sc.col_norm[j]; s += rd * rd; }
            err_rd = sc.c_scale * sqrt(s) / sc.norm_c_org;
            spmv(&A, x_bar, ax);
            s = 0.0;
            for (int i = 0; i < m; ++i) {
                double rp = fmax(fmin(AU[i] - ax[i], 0.0), AL[i] - ax[i]) * sc.row_norm[i]; s += rp * rp;
            }
            err_rp = sc.b_scale * sqrt(s) / sc.norm_b_org;
            if (iter == 0) {
                s = 0.0;
                for (int j = 0; j < n; ++j) {
                    double t = (x_bar[j] < l[j]) ? (l[j] - x_bar[j]) : ((x_bar[j] > u[j]) ? (x_bar[j] - u[j]) : 0.0);
                    x_tmp[j] = t / sc.col_norm[j]; s += x_tmp[j] * x_tmp[j];
                }
                err_rp = fmax(err_rp, sc.b_scale * sqrt(s));
            }
            kkt = fmax(fmax(err_rd, err_rp), gap);
            if (compute_gap) {
                double dp = 2.0 * g_dot;
                double wn = sigma * (lambda_max * g_dy) + g_dx / sigma + dp;
                if (wn < 0) {
                    lambda_max = -(dp + g_dx / sigma) / (sigma * g_dy) * 1.05;
                    wn = sqrt(-(dp + g_dx / sigma) * 0.05);
                } else wn = sqrt(wn);
                R.current_gap = wn;
            }
        }
        /* check_stopping, src/main_iterate.cu:406-420 */
        if (kkt < p->stop_tol) status = "OPTIMAL";
        else if (iter >= p->max_iter) status = "ITER_LIMIT";
        else if (elapsed > p->time_limit) status = "TIME_LIMIT";
        /* check_restart, src/main_iterate.cu:324-364 */
        R.restart_flag = 0;
        if (periodic) {
            if (R.first_restart) {
                if (iter == p->check_iter) {
                    R.first_restart = 0; R.restart_flag = 1; R.best_gap = R.current_gap; R.best_sigma = sigma;
                }
            } else {
                if (R.current_gap < 0) R.current_gap = 1e-6;
                if (R.current_gap <= 0.2 * R.last_gap) R.restart_flag = 1;
                if (R.current_gap <= 0.6 * R.last_gap && R.current_gap > 1.00 * R.save_gap) R.restart_flag = 2;
                if (R.inner >= 0.2 * iter) R.restart_flag = 3;
                if (R.best_gap > R.current_gap) { R.best_gap = R.current_gap; R.best_sigma = sigma; }
                R.save_gap = R.current_gap;
            }
        }
        if (strcmp(status, "CONTINUE") != 0) break;

        if (R.restart_flag > 0) {
            /* update_sigma, src/main_iterate.cu:367-404 */
            double pm = 0, dm = 0;
            for (int j = 0; j < n; ++j) { double d = x_bar[j] - x0[j]; pm += d * d; }
            for (int i = 0; i < m; ++i) { double d = y_bar[i] - y0[i]; dm += d * d; }
            pm = sqrt(pm); dm = sqrt(dm);
            if (pm > 1e-16 && dm > 1e-16 && pm < 1e12 && dm < 1e12) {
                double ratio = pm / dm / sqrt(lambda_max);
                double fact = exp(-0.05 * (R.current_gap / R.best_gap));
                double temp1 = fmax(fmin(err_rd, err_rp), fmin(gap, R.current_gap));
                double sig_c = exp(fact * log(ratio) + (1 - fact) * log(R.best_sigma));
                double kappa;
                if (temp1 > 9e-10) kappa = 1.0;
                else if (temp1 > 5e-10) kappa = fmax(fmin(sqrt(err_rd / err_rp), 100.0), 1e-2);
                else kappa = fmax(fmin(err_rd / err_rp, 100.0), 1e-2);
                sigma = kappa * sig_c;
            } else sigma = 1.0;
            /* do_restart, src/main_iterate.cu:312-322; upload_halpern_restart_params :54-66 */
            memcpy(x0, x_bar, sizeof(double) * n); memcpy(y0, y_bar, sizeof(double) * m);
            memcpy(x, x_bar, sizeof(double) * n); memcpy(y, y_bar, sizeof(double) * m);
            R.inner = 0; R.times += 1; R.save_gap = INFINITY; k_inner = 0;
        }
        /* one HPR iteration: src/cuda_kernels/HPR_cuda_kernels.cu:203-295 / 297-427 */
        {
            /* check iteration iff (iter+1)%check_iter==0 || restart || (iter+1)%step(iter+1)==0
             * (src/HPRLP.cu:295-296); only check iterations refresh x_bar,z_bar,x_tmp,y_bar,y_obj,y_tmp,
             * so ITER_LIMIT/TIME_LIMIT can return stale bars exactly like the reference (quirk #2). */
            const int chk = ((iter + 1) % p->check_iter == 0) || R.restart_flag > 0 ||
                            ((iter + 1) % step_of(iter + 1) == 0);
            const double f1 = 1.0 / (k_inner + 2.0), f2 = 1.0 - f1;
            const double lamsig = lambda_max * sigma, inv_lamsig = 1.0 / lamsig;
            spmv(&AT, y, w);
#pragma omp parallel for schedule(static)
            for (int j = 0; j < n; ++j) {
                double xi = x[j];
                double zt = fma(sigma, w[j] - c[j], xi);
                double xb = fmin(u[j], fmax(l[j], zt));
                double xh = 2.0 * xb - xi;
                x[j] = fma(f2, xh, f1 * x0[j]);
                x_hat[j] = xh;
                if (chk) { x_bar[j] = xb; z_bar[j] = (xb - zt) / sigma; x_tmp[j] = xb - xh; }
            }
            spmv(&A, x_hat, ax);
#pragma omp parallel for schedule(static)
            for (int i = 0; i < m; ++i) {
                double yi = y[i];
                double v = fma(-lamsig, yi, ax[i]);
                double d = fmax(AL[i] - v, fmin(AU[i] - v, 0.0));
                double yb = inv_lamsig * d;
                double yh = 2.0 * yb - yi;
                y[i] = fma(f2, yh, f1 * y0[i]);
                if (chk) { y_bar[i] = yb; y_obj[i] = v + d; y_tmp[i] = yb - yh; }
            }
            k_inner += 1;
        }
        for (int t = 0; t < n_trace; ++t)
            if (trace_iters[t] == iter + 1) {
                for (int j = 0; j < n; ++j) {
                    trace_x[(size_t)t * n + j] = sc.b_scale * (x_bar[j] / sc.col_norm[j]);
                    trace_z[(size_t)t * n + j] = sc.c_scale * (z_bar[j] * sc.col_norm[j]);
                }
                for (int i = 0; i < m; ++i) trace_y[(size_t)t * m + i] = sc.c_scale * (y_bar[i] / sc.row_norm[i]);
            }
        if (R.restart_flag > 0) {
            /* compute_weighted_norm, src/main_iterate.cu:486-515 */
            spmv(&A, x_tmp, ax);
            double dp = 2.0 * dot(ax, y_tmp, m), dy = dot(y_tmp, y_tmp, m), dx = dot(x_tmp, x_tmp, n);
            double wn = sigma * (lambda_max * dy) + dx / sigma + dp;
            if (wn < 0) { lambda_max = -(dp + dx / sigma) / (sigma * dy) * 1.05; wn = sqrt(-(dp + dx / sigma) * 0.05); }
            else wn = sqrt(wn);
            R.last_gap = wn;
        }
        R.inner += 1;
    }
    struct timespec t1; clock_gettime(CLOCK_MONOTONIC, &t1);
    /* collect_solution, src/utils.cu:143-200 */
    for (int j = 0; j < n; ++j) {
        x_out[j] = sc.b_scale * (x_bar[j] / sc.col_norm[j]);
        z_out[j] = sc.c_scale * (z_bar[j] * sc.col_norm[j]);
    }
    for (int i = 0; i < m; ++i) y_out[i] = sc.c_scale * (y_bar[i] / sc.row_norm[i]);
    if (info) {
        memset(info, 0, sizeof(*info));
        info->residuals = kkt; info->primal_obj = pobj; info->dual_obj = dobj; info->gap = gap;
        info->err_rp = err_rp; info->err_rd = err_rd; info->iter = iter;
        strncpy(info->status, status, 63);
        info->lambda_max = lambda_max; info->sigma = sigma; info->restarts = R.times;
        info->b_scale = sc.b_scale; info->c_scale = sc.c_scale; info->norm_b = sc.norm_b; info->norm_c = sc.norm_c;
        info->norm_b_org = sc.norm_b_org; info->norm_c_org = sc.norm_c_org; info->power_iters = piters;
        info->solve_seconds = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
    }
    free(A.rp); free(A.ci); free(A.v); free(AT.rp); free(AT.ci); free(AT.v);
    free(AL); free(AU); free(l); free(u); free(c); free(sc.row_norm); free(sc.col_norm);
    free(x); free(x0); free(x_hat); free(x_bar); free(z_bar); free(x_tmp);
    free(y); free(y0); free(y_bar); free(y_obj); free(y_tmp); free(w); free(ax);
    return 0;
}

/* Scaling only: returns the scaled problem so the GPU scaling kernels can be compared array by array. */
int oracle_scale(int m, int n, const int *rowPtr, const int *colIndex, double *values /* in/out */,
                 double *AL, double *AU, double *l, double *u, double *c, const oracle_params *p,
                 int *at_rowPtr, int *at_col, double *at_val, double *row_norm_out, double *col_norm_out,
                 double *scalars6 /* b_scale,c_scale,norm_b,norm_c,norm_b_org,norm_c_org */) {
    int nnz = rowPtr[m];
    csr_t A = {m, n, nnz, (int *)rowPtr, (int *)colIndex, values};
    csr_t AT = {n, m, nnz, at_rowPtr, at_col, at_val};
    oracle_transpose(m, n, nnz, rowPtr, colIndex, values, at_rowPtr, at_col, at_val);
    scaling_t sc; sc.row_norm = row_norm_out; sc.col_norm = col_norm_out;
    do_scaling(&A, &AT, AL, AU, c, l, u, p, &sc);
    scalars6[0] = sc.b_scale; scalars6[1] = sc.c_scale; scalars6[2] = sc.norm_b; scalars6[3] = sc.norm_c;
    scalars6[4] = sc.norm_b_org; scalars6[5] = sc.norm_c_org;
    return 0;
}

/* Power iteration only on a given (already scaled) matrix. Returns lambda (not yet x1.01). */
double oracle_power(int m, int n, const int *rowPtr, const int *colIndex, const double *values,
                    const double *z0, int max_iter, double tol, int *iters) {
    int nnz = rowPtr[m];
    csr_t A = {m, n, nnz, (int *)rowPtr, (int *)colIndex, (double *)values};
    csr_t AT = {n, m, nnz, (int *)malloc(sizeof(int) * (n + 1)), (int *)malloc(sizeof(int) * nnz),
                (double *)malloc(sizeof(double) * nnz)};
    oracle_transpose(m, n, nnz, rowPtr, colIndex, values, AT.rp, AT.ci, AT.v);
    double lam = power_method(&A, &AT, z0, max_iter, tol, iters);
    free(AT.rp); free(AT.ci); free(AT.v);
    return lam;
}

/* bench.py cpu_baseline leg: wall time of `iters` plain HPR iterations (x-phase + y-phase of
 * HPR_cuda_kernels.cu:297-427) on the given, unscaled problem with sigma = lambda = 1 -- a bounded
 * sample of the hot path on the host cores (OpenMP over rows).  Returns seconds. */
double oracle_time_iterations(int m, int n, const int *rowPtr, const int *colIndex, const double *values,
                              const double *AL, const double *AU, const double *l, const double *u, const double *c,
                              int iters, int *threads_out) {
    int nnz = rowPtr[m];
    csr_t A = {m, n, nnz, (int *)rowPtr, (int *)colIndex, (double *)values};
    csr_t AT = {n, m, nnz, (int *)malloc(sizeof(int) * (n + 1)), (int *)malloc(sizeof(int) * nnz),
                (double *)malloc(sizeof(double) * nnz)};
    oracle_transpose(m, n, nnz, rowPtr, colIndex, values, AT.rp, AT.ci, AT.v);
    VEC(x, n); VEC(x0, n); VEC(x_hat, n); VEC(y, m); VEC(y0, m); VEC(w, n); VEC(ax, m);
    int nthreads = 1;
#ifdef _OPENMP
#pragma omp parallel
    {
#pragma omp single
        nthreads = omp_get_num_threads();
    }
#endif
    if (threads_out) *threads_out = nthreads;
    const double sigma = 1.0, lamsig = 1.0, inv = 1.0;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int k = 0; k < iters; ++k) {
        const double f1 = 1.0 / (k + 2.0), f2 = 1.0 - f1;
        spmv(&AT, y, w);
#pragma omp parallel for schedule(static)
        for (int j = 0; j < n; ++j) {
            double xi = x[j], zt = fma(sigma, w[j] - c[j], xi);
            double xb = fmin(u[j], fmax(l[j], zt)), xh = 2.0 * xb - xi;
            x[j] = fma(f2, xh, f1 * x0[j]); x_hat[j] = xh;
        }
        spmv(&A, x_hat, ax);
#pragma omp parallel for schedule(static)
        for (int i = 0; i < m; ++i) {
            double yi = y[i], v = fma(-lamsig, yi, ax[i]);
            double d = fmax(AL[i] - v, fmin(AU[i] - v, 0.0));
            double yb = inv * d, yh = 2.0 * yb - yi;
            y[i] = fma(f2, yh, f1 * y0[i]);
        }
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free(AT.rp); free(AT.ci); free(AT.v);
    free(x); free(x0); free(x_hat); free(y); free(y0); free(w); free(ax);
    return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}
