/* batched_solver.h -- the batched shared-A entry points of the C ABI.
 *
 * Drop-in for the reference's include/batched_solver.h:23-33 (same two symbols, same argument order and types, struct
 * returned by value); implemented by hpr-lp-c_b200/csrc/batched.cu (reference src/batched_solver.cu:939-1105).
 *
 *   solve_batched          B = batch_size LPs that share the sparse matrix A of `model`; instance k is
 *                              min  C[:,k]' x + obj_constants[k]   s.t.  AL[:,k] <= A x <= AU[:,k],  l[:,k] <= x <= u[:,k]
 *                          C, l, u: n x B and AL, AU: m x B, column-major (instance k is one contiguous column, i.e. a
 *                          C-ordered (B, n) / (B, m) array); obj_constants (B entries) and param may be NULL.
 *   free_batched_results   releases every array a result owns (x, y, z, primal_obj, residuals, gap, iter, status).
 */
#pragma once
#ifndef HPRLP_B200_BATCHED_ENTRY_POINTS
#define HPRLP_B200_BATCHED_ENTRY_POINTS
#include "structs.h"
#ifdef __cplusplus
extern "C" {
#endif
HPRLP_batched_results solve_batched(const LP_info_cpu *model, int batch_size, const HPRLP_FLOAT *C, const HPRLP_FLOAT *AL,
                                    const HPRLP_FLOAT *AU, const HPRLP_FLOAT *l, const HPRLP_FLOAT *u,
                                    const HPRLP_FLOAT *obj_constants, const HPRLP_parameters *param);
void free_batched_results(HPRLP_batched_results *results);
#ifdef __cplusplus
} /* extern "C" */
#endif
#endif /* HPRLP_B200_BATCHED_ENTRY_POINTS */
