// abi_guard.h -- no C++ exception may cross the extern "C" boundary: ctypes, Julia ccall and MEX callers would see
// std::terminate.  Every entry point runs its body through one of these guards: the error is printed as "[error] ..."
// (the reference's own convention, src/HPRLP.cu:329-337) and an "ERROR" result / null handle / -1 is returned.
#pragma once
#include <cstdio>
#include <cstring>
#include <exception>

#include <cuda_runtime.h>

#include "../../include/structs.h"

namespace hpr {

inline HPRLP_results abi_error_result() {   // reference src/HPRLP.cu:66-79
    HPRLP_results r;
    std::memset(r.status, 0, sizeof(r.status));
    std::strncpy(r.status, "ERROR", sizeof(r.status) - 1);
    r.iter = 0; r.time = 0.0; r.primal_obj = 0.0; r.residuals = 0.0; r.gap = 0.0;
    r.x = nullptr; r.y = nullptr; r.z = nullptr;
    return r;
}

inline void abi_report(const char *where, const char *what) {
    std::fprintf(stderr, "[error] %s: %s\n", where, what);
    std::fflush(stderr);
}

template <class R, class F, class E>
inline R abi_guard(const char *where, F &&body, E &&on_error) {
    try {
        return body();
    } catch (const std::exception &e) {
        abi_report(where, e.what());
    } catch (...) {
        abi_report(where, "unknown exception");
    }
    cudaGetLastError();   // a recoverable CUDA error must not be found again by the next call's cudaGetLastError check
    return on_error();
}

template <class F>
inline HPRLP_results abi_guard_results(const char *where, F &&body) {
    return abi_guard<HPRLP_results>(where, body, [] { return abi_error_result(); });
}
template <class F>
inline int abi_guard_int(const char *where, F &&body) {
    return abi_guard<int>(where, body, [] { return -1; });
}
template <class F>
inline double abi_guard_double(const char *where, F &&body) {
    return abi_guard<double>(where, body, [] { return -1.0; });
}

}  // namespace hpr
