#include "nccl_shim.h"

#include <dlfcn.h>

#include <mutex>
#include <stdexcept>
#include <string>

namespace hpr {

const NcclApi &nccl() {
    static NcclApi api;
    static std::once_flag once;
    static std::string error;
    std::call_once(once, []() {
        void *h = nullptr;
        // an NCCL already mapped into the process (PyTorch's) is reused; otherwise the system library
        for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
            h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (h) break;
        }
        if (!h) { error = std::string("cannot load libnccl.so.2: ") + dlerror(); return; }
        auto sym = [&](const char *n) { void *p = dlsym(h, n); if (!p) error = std::string("missing NCCL symbol ") + n; return p; };
        api.GetUniqueId = reinterpret_cast<int (*)(NcclUniqueId *)>(sym("ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<int (*)(NcclComm *, int, NcclUniqueId, int)>(sym("ncclCommInitRank"));
        api.CommInitAll = reinterpret_cast<int (*)(NcclComm *, int, const int *)>(sym("ncclCommInitAll"));
        api.CommDestroy = reinterpret_cast<int (*)(NcclComm)>(sym("ncclCommDestroy"));
        api.CommAbort = reinterpret_cast<int (*)(NcclComm)>(sym("ncclCommAbort"));
        api.ReduceScatter = reinterpret_cast<int (*)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t)>(sym("ncclReduceScatter"));
        api.AllGather = reinterpret_cast<int (*)(const void *, void *, size_t, int, NcclComm, cudaStream_t)>(sym("ncclAllGather"));
        api.AllReduce = reinterpret_cast<int (*)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t)>(sym("ncclAllReduce"));
        api.GroupStart = reinterpret_cast<int (*)()>(sym("ncclGroupStart"));
        api.GroupEnd = reinterpret_cast<int (*)()>(sym("ncclGroupEnd"));
        api.GetErrorString = reinterpret_cast<const char *(*)(int)>(sym("ncclGetErrorString"));
    });
    if (!error.empty()) throw std::runtime_error("NCCL unavailable: " + error);
    return api;
}

}  // namespace hpr
