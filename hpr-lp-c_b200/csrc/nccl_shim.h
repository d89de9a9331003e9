// nccl_shim.h -- NCCL loaded lazily with dlopen (only the row-partitioned multi-GPU path needs it), so that
// libhprlp.so carries no DT_NEEDED on libnccl and never clashes with another NCCL already loaded in the
// process (e.g. the one bundled with PyTorch).  Only the handful of entry points used are bound.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace hpr {

struct NcclUniqueId { char internal[128]; };
typedef struct ncclComm *NcclComm;

struct NcclApi {
    int (*GetUniqueId)(NcclUniqueId *) = nullptr;
    int (*CommInitRank)(NcclComm *, int, NcclUniqueId, int) = nullptr;
    int (*CommInitAll)(NcclComm *, int, const int *) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*CommAbort)(NcclComm) = nullptr;
    int (*ReduceScatter)(const void *, void *, size_t /*recvcount*/, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t /*sendcount*/, int, NcclComm, cudaStream_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int /*dtype*/, int /*op*/, NcclComm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    const char *(*GetVersionDummy)() = nullptr;
};

enum { kNcclSum = 0, kNcclMax = 2, kNcclFloat64 = 8 };

// Throws std::runtime_error when libnccl cannot be loaded (no silent fallback).
const NcclApi &nccl();

}  // namespace hpr
