// presolve.cpp -- bridge to the PSLP presolver (third-party, Apache-2.0, v0.0.8; vendored by the
// reference under third_party/PSLP and driven from src/pslp_integration.cpp).  Host-side and OUT OF
// SCOPE for the rebuild: the presolver is an external dependency that is
// linked when the build finds its sources (HPRLP_WITH_PSLP, see Makefile), never copied into this
// repository.  Without it use_presolve=true solves the original model, which is also the
// reference's own fallback when presolve fails (src/pslp_integration.cpp:677-691).
#include <cstdio>

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
#endif
