// batched.cu -- solve_batched / free_batched_results (reference src/batched_solver.cu:939-1105).
#include <cstdlib>
#include <cstring>
#include <algorithm>

#include "../../include/batched_solver.h"
#include "engine.h"

namespace {
HPRLP_batched_results make_batched_error(const char *status, int m, int n, int B) {   // reference :356-368
    HPRLP_batched_results r;
    r.m = m; r.n = n; r.batch_size = B;
    if (B > 0) {
        r.status = static_cast<char *>(std::calloc(static_cast<size_t>(B) * 64, sizeof(char)));
        for (int k = 0; k < B; ++k) std::strncpy(r.status + 64 * k, status, 63);
    }
    return r;
}
}  // namespace

extern "C" HPRLP_batched_results solve_batched(const LP_info_cpu *model, int batch_size, const HPRLP_FLOAT *C,
                                                const HPRLP_FLOAT *AL, const HPRLP_FLOAT *AU, const HPRLP_FLOAT *l,
                                                const HPRLP_FLOAT *u, const HPRLP_FLOAT *obj_constants,
                                                const HPRLP_parameters *param) {
    (void)obj_constants; (void)param;
    if (!model || !model->A || batch_size <= 0 || !C || !AL || !AU || !l || !u) {
        return make_batched_error("ERROR", model ? model->m : 0, model ? model->n : 0, std::max(batch_size, 0));
    }
    return make_batched_error("ERROR", model->m, model->n, batch_size);
}

extern "C" void free_batched_results(HPRLP_batched_results *results) {   // reference :1094-1105
    if (!results) return;
    std::free(results->x); std::free(results->y); std::free(results->z);
    std::free(results->primal_obj); std::free(results->residuals); std::free(results->gap);
    std::free(results->iter); std::free(results->status);
    *results = HPRLP_batched_results{};
}
