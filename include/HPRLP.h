#ifndef HPRLP_H
#define HPRLP_H
/*
 * HPR-LP C API, B200-native engine.  Same seven extern "C" entry points, argument meaning,
 * ownership and error behaviour as the reference PolyU-IOR/HPR-LP-C `include/HPRLP.h`
 * (file:line of the declaration each one replaces is given below), so the reference's
 * Python/Julia/MATLAB bindings, examples and build/solve_mps_file link unchanged.
 *
 *   min c'x   s.t.  AL <= A x <= AU,   l <= x <= u      (A: m x n CSR, fp64, int32 indices)
 */
#include "structs.h"
#include "batched_solver.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Core solve on an already built model, no presolve (reference include/HPRLP.h:41,
 * src/HPRLP.cu:116-311). */
HPRLP_results HPRLP_main_solve(const LP_info_cpu *lp_info_cpu, const HPRLP_parameters *param);

/* Build a model from CSR (or CSC when is_csc) arrays; inputs are copied.  Returns NULL and
 * prints "[error] ..." on stderr for m<=0, n<=0, nnz<=0 or a NULL array
 * (reference include/HPRLP.h:105-111, src/HPRLP.cu:321-446). */
LP_info_cpu* create_model_from_arrays(int m, int n, int nnz,
                                      const int *rowPtr, const int *colIndex,
                                      const HPRLP_FLOAT *values,
                                      const HPRLP_FLOAT *AL, const HPRLP_FLOAT *AU,
                                      const HPRLP_FLOAT *l, const HPRLP_FLOAT *u,
                                      const HPRLP_FLOAT *c,
                                      bool is_csc = false);

/* Build a model from a free-format MPS file (.mps or .mps.gz)
 * (reference include/HPRLP.h:140, src/HPRLP.cu:451-488). */
LP_info_cpu* create_model_from_mps(const char* mps_file_path);

/* Solve; param may be NULL (defaults).  result.x/y/z are malloc'd, caller frees
 * (reference include/HPRLP.h:180, src/HPRLP.cu:493-524). */
HPRLP_results solve(const LP_info_cpu *model, const HPRLP_parameters *param);

/* Free a model; NULL is a no-op (reference include/HPRLP.h:202, src/HPRLP.cu:529-537). */
void free_model(LP_info_cpu *model);

#ifdef __cplusplus
}
#endif

#endif
