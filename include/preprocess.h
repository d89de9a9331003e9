#ifndef HPRLP_COMPAT_PREPROCESS_H
#define HPRLP_COMPAT_PREPROCESS_H
/* Compatibility header: the reference's bindings include "preprocess.h" (e.g.
 * bindings/python/src/hprlp_pybind.cpp:18-20, bindings/matlab/src/hprlp_mex.cpp:11-13) but use
 * only the seven extern "C" symbols of HPRLP.h; the engine internals are private here. */
#include "structs.h"
#endif
