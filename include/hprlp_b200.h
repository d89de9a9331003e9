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
    int reserved0;                  /* ranks of the row partition (1 = single GPU) */
    int peer_exchange;              /* 1: per-iteration exchange = our fused kernel over NVLink peer memory; 0: NCCL collectives */
    int reserved1;
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
/* Average device time (ms) of one launch: which = 0 fused x-phase (A^T pass), 1 fused y-phase (A pass); row-partitioned
 * engines: 2 partial A_p^T y_p pass, 3 x-update on the owned x-block, 4 reduce-scatter + all-gather pair (collective:
 * every rank calls it).  Timing only: the iterates are garbage afterwards. */
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

/* solve_batched for dense inputs held ROW-major (layout 1: C-ordered (n, B) / (m, B) arrays, element (i, k) at i*B + k),
 * as the reference's Python API receives them; layout 0 = column-major = solve_batched.  The arrays are uploaded as they
 * are and re-laid out on the device: no host-side re-packing (reference bindings/python/src/hprlp_pybind.cpp:343-356,
 * 413-455 copies all five arrays element by element before every call).  Outputs column-major as in solve_batched. */
HPRLP_batched_results hprlp_b200_solve_batched_layout(const LP_info_cpu *model, int batch_size, const HPRLP_FLOAT *C,
                                                      const HPRLP_FLOAT *AL, const HPRLP_FLOAT *AU, const HPRLP_FLOAT *l,
                                                      const HPRLP_FLOAT *u, const HPRLP_FLOAT *obj_constants,
                                                      const HPRLP_parameters *param, int layout);

/* One LP row-block partitioned over n_gpus GPUs of one node (devices param->device_number ...).  GPU p owns a block of
 * rows of A (+ its transpose, y-side vectors) and the x-block J_p.  Per iteration: partial A_p^T y_p -> NCCL
 * reduce-scatter over NVLink -> x-update on J_p -> NCCL all-gather of x_hat -> fused y-phase on the local rows;
 * <= 9 residual scalars all-reduced per check.  Same results struct as solve() (x, z: n; y: m); no presolve.
 * n_gpus is clamped to the visible devices; 1 GPU = HPRLP_main_solve.  One host thread per GPU (ncclCommInitAll).
 * (new functionality; the reference is single-GPU, src/HPRLP.cu:51-64) */
HPRLP_results hprlp_b200_solve_partitioned(const LP_info_cpu *model, const HPRLP_parameters *param, int n_gpus,
                                           int quiet, hprlp_b200_info *info);

/* The same partitioned engine with n_ranks LOGICAL ranks that all live on device param->device_number; the exchanges
 * are plain kernels between host barriers.  Parity-test path for single-GPU boxes (not for speed). */
HPRLP_results hprlp_b200_solve_partitioned_local(const LP_info_cpu *model, const HPRLP_parameters *param, int n_ranks,
                                                 int quiet, hprlp_b200_info *info);

/* One process per GPU (torchrun / MPI style launch).  Rank 0 calls hprlp_b200_nccl_unique_id and distributes the 128
 * bytes (bench.py: torch.distributed broadcast); every rank creates ONE communicator handle for device `device` and
 * reuses it for any number of partitioned solves (the NCCL bootstrap costs seconds, a solve tenths of a second).
 * Every rank passes the FULL host model; the rank uploads and owns its row block.  All ranks return the full solution.
 * hprlp_b200_engine_create_rank is the resident-engine variant (hprlp_b200_engine_run etc. are then collective: every
 * rank calls them with the same arguments).  A communicator must outlive the engines created on it. */
typedef struct hprlp_b200_comm hprlp_b200_comm;
int hprlp_b200_nccl_unique_id(char *out128);
hprlp_b200_comm *hprlp_b200_comm_create(const char *uid128, int rank, int nranks, int device);
void hprlp_b200_comm_destroy(hprlp_b200_comm *comm);
HPRLP_results hprlp_b200_solve_partitioned_rank(const LP_info_cpu *model, const HPRLP_parameters *param, hprlp_b200_comm *comm,
                                                int quiet, hprlp_b200_info *info);
hprlp_b200_engine *hprlp_b200_engine_create_rank(const LP_info_cpu *model, const HPRLP_parameters *param, hprlp_b200_comm *comm);

/* Diagnostic: ms of one in-place reduce-scatter + all-gather pair (out_ms[0]) and of one all-reduce (out_ms[1]) of
 * `count` doubles over n_gpus GPUs, issued exactly as the partitioned solver issues them.  Returns 0 on success. */
int hprlp_b200_nccl_exchange_ms(int n_gpus, long long count, int reps, double *out_ms);

/* The synthetic "uniform" LP of BASELINE.json generated shard by shard ON the GPUs (same counter-based generator as
 * tools/synth_lp.c, bit-identical matrix) and solved row-partitioned: the path for instances whose CSR + transpose
 * exceed one GPU and the int32 host ABI (config 5, nnz = 6e9).  Each shard needs < 2^31 nonzeros.  *obj_star = the
 * constructed optimal value; want_solution == 0 returns no x/y/z. */
HPRLP_results hprlp_b200_solve_partitioned_synth(long long m, int n, int K, unsigned long long seed,
                                                 const HPRLP_parameters *param, int n_gpus, int quiet, int want_solution,
                                                 double *obj_star, hprlp_b200_info *info);
/* The same with one process per GPU (see hprlp_b200_solve_partitioned_rank): this process generates row block `rank`. */
HPRLP_results hprlp_b200_solve_partitioned_synth_rank(long long m, int n, int K, unsigned long long seed,
                                                      const HPRLP_parameters *param, hprlp_b200_comm *comm,
                                                      int quiet, int want_solution, double *obj_star, hprlp_b200_info *info);
/* Test hook: rows [row0, row0+rows) of that matrix, generated on the device, copied to host (rows*K entries each). */
int hprlp_b200_synth_rows(int n, int K, unsigned long long seed, long long row0, int rows, int *col_out, double *val_out);

/* Presolve step only (PSLP bridge, host): fills *reduced and *handle, returns 1 on success, 0 when presolve is
 * unavailable or failed (the caller then solves the original model, reference src/HPRLP.cu:508-511).
 * hprlp_b200_presolve_free releases both. */
int hprlp_b200_presolve(const LP_info_cpu *model, const HPRLP_parameters *param, LP_info_cpu *reduced, void **handle);
void hprlp_b200_presolve_free(void *handle, LP_info_cpu *reduced);

/* Finished solves keep their device arena cached in a private stream-ordered memory pool (per device, bounded by
 * HPRLP_POOL_RETAIN_MB, default 8192) so that repeated solve() calls skip cudaMalloc/cudaFree.  This call returns all
 * cached memory to the driver (cudaMemPoolTrimTo 0).  The reference frees everything at the end of each solve. */
void hprlp_b200_release_cached_memory(void);

/* First solve of a process: CUDA context creation, module load and cuRAND's first generator cost 1-3 s, more than a
 * whole configs[1] solve.  This call starts them on a background thread and returns immediately, so they overlap host
 * work that precedes the solve (MPS parsing in build/solve_mps_file; the PSLP presolve inside solve(), which starts the
 * warm-up itself).  The next solve on `device` joins the thread.  HPRLP_NO_WARMUP=1 disables it. */
void hprlp_b200_warmup(int device);

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
