#ifndef HPRLP_STRUCTS_H
#define HPRLP_STRUCTS_H
/*
 * Public POD types of the HPR-LP C API, B200-native engine.
 *
 * Drop-in boundary: the layouts below are byte-identical (x86-64) to the reference
 * PolyU-IOR/HPR-LP-C `include/structs.h`:
 *   sparseMatrix          40 B  (reference include/structs.h:16-22)
 *   HPRLP_parameters      40 B  (reference include/structs.h:25-40)
 *   HPRLP_results        160 B  (reference include/structs.h:44-65)
 *   HPRLP_batched_results 112 B (reference include/structs.h:68-90)
 *   LP_info_cpu           64 B  (reference include/structs.h:231-240)
 * `tests/test_abi.py` checks every offset with a compiled probe.  The reference also
 * declares its internal GPU workspace (cuSPARSE/cuBLAS handles) in this header; that is
 * not part of the ABI and lives in hpr-lp-c_b200/csrc/engine.h here, so consumers of this
 * header need neither cuBLAS nor cuSPARSE headers.
 *
 * Source compatibility: consumers of the reference header also get <cmath> (INFINITY), <limits>, <string> and
 * <vector> through it (the reference pulls them in next to cublas_v2.h); its own examples rely on that
 * (examples/cpp/example_direct_lp.cpp uses INFINITY with no include of its own), so they are included here too.
 */
#include <math.h>
#include <stdint.h>
#ifdef __cplusplus
#include <cmath>
#include <limits>
#include <string>
#include <vector>
#endif

#define HPRLP_FLOAT double

/* CSR matrix, int32 indices (reference include/structs.h:16-22). */
struct sparseMatrix {
    int row, col;
    int numElements;
    int *colIndex;
    int *rowPtr;
    HPRLP_FLOAT *value;
};

/* Solver parameters (reference include/structs.h:25-40).  CUSPARSE_spmv and autotune_verbose
 * are accepted and ignored: this engine has a single hand-written backend. */
struct HPRLP_parameters {
    int max_iter = INT32_MAX;
    HPRLP_FLOAT stop_tol = 1e-4;
    HPRLP_FLOAT time_limit = 3600.0;
    int device_number = 0;
    int check_iter = 150;
    bool CUSPARSE_spmv = false;
    bool autotune_verbose = false;

    bool use_CR_scaling = true;
    bool use_Ruiz_scaling = true;
    bool use_Pock_Chambolle_scaling = true;
    bool use_bc_scaling = true;
    bool use_presolve = true;
};

/* Result of a single solve (reference include/structs.h:44-65); x/y/z are malloc'd, the
 * caller frees them with free(). */
struct HPRLP_results {
    HPRLP_FLOAT residuals;
    HPRLP_FLOAT primal_obj;
    HPRLP_FLOAT gap;

    HPRLP_FLOAT time4 = 0.0;
    HPRLP_FLOAT time6 = 0.0;
    HPRLP_FLOAT time8 = 0.0;
    HPRLP_FLOAT time = 0.0;
    int iter4 = 0;
    int iter6 = 0;
    int iter8 = 0;
    int iter = 0;

    char status[64]; /* "OPTIMAL", "TIME_LIMIT", "ITER_LIMIT", "ERROR" */

    HPRLP_FLOAT *x = nullptr;
    HPRLP_FLOAT *y = nullptr;
    HPRLP_FLOAT *z = nullptr;
};

/* Result of solve_batched (reference include/structs.h:68-90). Column-major host arrays:
 * x/z are n x batch_size, y is m x batch_size; status is batch_size 64-byte slots. */
struct HPRLP_batched_results {
    int m = 0;
    int n = 0;
    int batch_size = 0;

    HPRLP_FLOAT *x = nullptr;
    HPRLP_FLOAT *y = nullptr;
    HPRLP_FLOAT *z = nullptr;

    HPRLP_FLOAT *primal_obj = nullptr;
    HPRLP_FLOAT *residuals = nullptr;
    HPRLP_FLOAT *gap = nullptr;
    int *iter = nullptr;

    char *status = nullptr;

    HPRLP_FLOAT time = 0.0;
    HPRLP_FLOAT setup_time = 0.0;
    HPRLP_FLOAT solve_time = 0.0;
    HPRLP_FLOAT power_time = 0.0;
};

/* Host model (reference include/structs.h:231-240). */
struct LP_info_cpu {
    int m, n;
    sparseMatrix *A;
    HPRLP_FLOAT *AL;
    HPRLP_FLOAT *AU;
    HPRLP_FLOAT *c;
    HPRLP_FLOAT *l;
    HPRLP_FLOAT *u;
    HPRLP_FLOAT obj_constant;
};

/* Convenience array bundle (reference include/structs.h:285-307). */
struct HPRLP_LP_Data {
    int m;
    int n;
    int nnz;
    int *rowPtr;
    int *colIndex;
    HPRLP_FLOAT *values;
    bool is_csc;
    HPRLP_FLOAT *AL;
    HPRLP_FLOAT *AU;
    HPRLP_FLOAT *l;
    HPRLP_FLOAT *u;
    HPRLP_FLOAT *c;
};

#endif
