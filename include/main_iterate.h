#ifndef HPRLP_COMPAT_MAIN_ITERATE_H
#define HPRLP_COMPAT_MAIN_ITERATE_H
/* Compatibility header: the reference's bindings include "main_iterate.h" (e.g.
 * bindings/python/src/hprlp_pybind.cpp:18-20, bindings/matlab/src/hprlp_mex.cpp:11-13) but use
 * only the seven extern "C" symbols of HPRLP.h; the engine internals are private here. */
#include "structs.h"
#endif
