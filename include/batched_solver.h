#ifndef HPRLP_BATCHED_SOLVER_H
#define HPRLP_BATCHED_SOLVER_H
/* Batched shared-A entry points; replaces reference include/batched_solver.h:23-33. */
#include "structs.h"

#ifdef __cplusplus
extern "C" {
#endif

/*
 * Solve batch_size LPs sharing the sparse matrix A of `model`:
 *   min c_k'x + obj_constants[k]  s.t.  AL_k <= A x <= AU_k,  l_k <= x <= u_k.
 * Dense inputs are column-major: C, l, u are n x batch_size; AL, AU are m x batch_size.
 * obj_constants and param may be NULL.  (reference src/batched_solver.cu:939-1092)
 */
HPRLP_batched_results solve_batched(const LP_info_cpu *model,
                                    int batch_size,
                                    const HPRLP_FLOAT *C,
                                    const HPRLP_FLOAT *AL,
                                    const HPRLP_FLOAT *AU,
                                    const HPRLP_FLOAT *l,
                                    const HPRLP_FLOAT *u,
                                    const HPRLP_FLOAT *obj_constants,
                                    const HPRLP_parameters *param);

/* Frees every array of a batched result (reference src/batched_solver.cu:1094-1105). */
void free_batched_results(HPRLP_batched_results *results);

#ifdef __cplusplus
}
#endif

#endif
