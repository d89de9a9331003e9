#ifndef HPRLP_B200_H
#define HPRLP_B200_H
/*
 * Extended C ABI of the B200-native engine (additions, not replacements): step-wise access to
 * the same engine that backs solve()/HPRLP_main_solve(), used by tests/ and bench.py to
 * (a) feed a fixed power-iteration start vector, (b) snapshot iterates for the 1e-10
 * iterate-parity check, (c) time the hot path with inputs resident in HBM, (d) time a single
 * fused kernel for the roofline.  Plain pointers and sizes only.
 */
#include "structs.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hprlp_b200_engine hprlp_b200_engine;

typedef struct {
    double lambda_max, sigma;
    double setup_seconds, scaling_seconds, power_seconds;
    double loop_device_ms;          /* CUDA-event time of the main loop on the engine stream */
    int restarts, power_iters;
    long long kernel_launches;      /* engine kernels launched by this call (graph nodes counted) */
    double b_scale, c_scale, norm_b, norm_c, norm_b_org, norm_c_org;
    int lanes_A, lanes_AT;          /* per-matrix load-balance choice (lanes per row) */
    int items_A, items_AT;
    int bands_A;                    /* column bands of A (0 = single pass; >1 when n doubles exceed the L2 budget) */
    int reserved0;
} hprlp_b200_info;

/* HPRLP_main_solve with hooks.  power_z0 (host, length m) overrides the cuRAND start vector when
 * non-NULL.  trace_*: after the check iteration with iter+1 == trace_iters[t] the unscaled
 * (x_bar,y_bar,z_bar) -- what solve() returns for max_iter = trace_iters[t] -- are stored in row t.
 * quiet != 0 suppresses the log.  (reference src/HPRLP.cu:116-311) */
HPRLP_results hprlp_b200_solve_ex(const LP_info_cpu *model, const HPRLP_parameters *param,
                                  const double *power_z0, int n_trace, const int *trace_iters,
                                  double *trace_x, double *trace_y, double *trace_z, int quiet,
                                  hprlp_b200_info *info);

/* The power-iteration start vector the engine uses for m rows (cuRAND XORWOW seed 1, N(0,1)+1e-8;
 * 1e-8 for odd m): reference src/power_iteration.cu:44-57.  out: host, length m. */
int hprlp_b200_power_start(int m, int device, double *out);

/* Resident engine: upload + scaling + power iteration done once; hprlp_b200_engine_run then advances
 * the full HPR driver (checks, restarts, sigma updates included) by `iters` iterations with no
 * host<->device traffic except the 9-scalar residual fetches. */
hprlp_b200_engine *hprlp_b200_engine_create(const LP_info_cpu *model, const HPRLP_parameters *param);
/* Runs `iters` more HPR iterations; returns the device time in ms (CUDA events on the engine stream). */
double hprlp_b200_engine_run(hprlp_b200_engine *e, int iters);
/* Average device time (ms) of one fused kernel launch: which = 0 x-phase (A^T pass), 1 y-phase (A pass). */
double hprlp_b200_engine_time_phase(hprlp_b200_engine *e, int which, int reps);
/* Current KKT residual / objective of the resident engine (one residual pass). */
int hprlp_b200_engine_residuals(hprlp_b200_engine *e, double *kkt, double *primal_obj, double *dual_obj);
void hprlp_b200_engine_info(hprlp_b200_engine *e, hprlp_b200_info *info);
void hprlp_b200_engine_destroy(hprlp_b200_engine *e);

/* Scaling only (reference src/scaling.cu:88-216): returns scaled A values (CSR order of the model),
 * scaled A^T values and indices, the scaled vectors and the accumulated norms, all on the host. */
int hprlp_b200_scale_only(const LP_info_cpu *model, const HPRLP_parameters *param,
                          double *A_val, int *AT_rowPtr, int *AT_col, double *AT_val,
                          double *AL, double *AU, double *l, double *u, double *c,
                          double *row_norm, double *col_norm, double *scalars6);

/* solve_batched sharded over n_gpus GPUs of one node (devices param->device_number ... +n_gpus-1): contiguous
 * instance shards, A replicated, no per-iteration collective.  solve_batched itself calls this with
 * n_gpus = $HPRLP_NUM_GPUS (default 1).  (new functionality; reference is single-GPU) */
HPRLP_batched_results hprlp_b200_solve_batched_multi(const LP_info_cpu *model, int batch_size, const HPRLP_FLOAT *C,
                                                     const HPRLP_FLOAT *AL, const HPRLP_FLOAT *AU, const HPRLP_FLOAT *l,
                                                     const HPRLP_FLOAT *u, const HPRLP_FLOAT *obj_constants,
                                                     const HPRLP_parameters *param, int n_gpus);

/* One LP row-block partitioned over n_gpus GPUs of one node (devices param->device_number ...): NCCL all-reduce of
 * A^T y per iteration over NVLink, fused y-phase on the local rows, all-reduce of <= 4 residual scalars per check.
 * Same results struct as solve(); no presolve.  n_gpus is clamped to the visible devices; 1 GPU = HPRLP_main_solve.
 * (new functionality; the reference is single-GPU) */
HPRLP_results hprlp_b200_solve_partitioned(const LP_info_cpu *model, const HPRLP_parameters *param, int n_gpus,
                                           int quiet, hprlp_b200_info *info);

/* The synthetic "uniform" LP of BASELINE.json generated shard by shard ON the GPUs (same counter-based generator as
 * tools/synth_lp.c, bit-identical matrix) and solved row-partitioned: the path for instances whose CSR + transpose
 * exceed one GPU and the int32 host ABI (config 5, nnz = 6e9).  Each shard needs < 2^31 nonzeros.  *obj_star = the
 * constructed optimal value; want_solution == 0 returns no x/y/z. */
HPRLP_results hprlp_b200_solve_partitioned_synth(long long m, int n, int K, unsigned long long seed,
                                                 const HPRLP_parameters *param, int n_gpus, int quiet, int want_solution,
                                                 double *obj_star, hprlp_b200_info *info);
/* Test hook: rows [row0, row0+rows) of that matrix, generated on the device, copied to host (rows*K entries each). */
int hprlp_b200_synth_rows(int n, int K, unsigned long long seed, long long row0, int rows, int *col_out, double *val_out);

/* Presolve step only (PSLP bridge, host): fills *reduced and *handle, returns 1 on success, 0 when presolve is
 * unavailable or failed (the caller then solves the original model, reference src/HPRLP.cu:508-511).
 * hprlp_b200_presolve_free releases both. */
int hprlp_b200_presolve(const LP_info_cpu *model, const HPRLP_parameters *param, LP_info_cpu *reduced, void **handle);
void hprlp_b200_presolve_free(void *handle, LP_info_cpu *reduced);

/* cudaProfilerStart/Stop of the library's (statically linked) CUDA runtime: lets `ncu --profile-from-start off`
 * capture only the timed region of bench.py. */
void hprlp_b200_profiler_start(void);
void hprlp_b200_profiler_stop(void);

/* Library identification: returns "hprlp-b200 <engine string>". */
const char *hprlp_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif
