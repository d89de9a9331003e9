// collective.h -- the exchange steps of the row-partitioned mode: GPU p owns a row block of A and the
// x-block J_p.  Per HPR iteration: reduce-scatter of the partial A_p^T y_p (every GPU ends up with the column sums of
// ITS x-block), x-update on n/P entries, all-gather of the x_hat blocks; <= 9 residual scalars are all-reduced per
// check.  Two transports behind one interface:
//   * NCCL over NVLink 5 / NVSwitch (ncclReduceScatter / ncclAllGather / ncclAllReduce, in place), one communicator
//     per GPU -- created with ncclCommInitAll (one host thread per GPU) or ncclCommInitRank (one process per GPU,
//     unique id distributed by the caller, e.g. torch.distributed under torchrun);
//   * "local": P logical ranks that can address each other's buffers directly (all on ONE device, or peer access
//     enabled), synchronised by host barriers, reduced by plain kernels in rank order.  It exists so that the whole
//     partitioned code path (row blocks, x-block ownership, every exchange) is parity-tested on a 1-GPU box:
//     collectives of ranks that share a GPU must not wait on each other on the device (B200_PROFILING.md), so the
//     waiting is done by the host threads.
// The reference has no counterpart (single GPU, src/HPRLP.cu:51-64).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include "nccl_shim.h"

namespace hpr {

class Collective {
   public:
    int nranks = 1, rank = 0;
    virtual ~Collective() {}
    // in-place sum (or max) over ranks of buf[0..count)
    virtual void all_reduce(double *buf, size_t count, bool max_op, cudaStream_t st) = 0;
    // buf = nranks blocks of `block` doubles.  On return block `rank` holds the sum over ranks of that block
    // (the other blocks are unspecified).
    virtual void reduce_scatter_inplace(double *buf, size_t block, cudaStream_t st) = 0;
    // buf = nranks blocks; block `rank` is this rank's contribution; on return every block is filled.
    virtual void all_gather_inplace(double *buf, size_t block, cudaStream_t st) = 0;
    // Unblocks peers stuck in a collective after this rank failed (ncclCommAbort / poisoned barrier).
    virtual void abort() {}
    virtual const char *name() const = 0;
};

Collective *make_nccl_collective(NcclComm comm, int nranks, int rank);   // takes ownership of the communicator

// ---------------------------------------------------------------------------------------------------------------
// Peer-memory exchange (NVLink 5 / NVSwitch P2P): the per-iteration exchange of the row-partitioned mode as ONE kernel
// of ours instead of two NCCL collectives.  Every rank exposes three cudaMalloc'ed buffers to all the others (raw
// pointers + cudaDeviceEnablePeerAccess inside one process, CUDA IPC handles between processes; the bootstrap runs
// over the rank's NCCL communicator):
//   w      npad doubles = one receive slot of `xblock` doubles per rank: slot p holds rank p's partial A_p^T y_p restricted
//          to THIS rank's x-block, stored there by rank p's own SpMV pass (SpmvPushOp: the reduce-scatter is fused into
//          the pass, its NVLink traffic runs under the pass)
//   xhat   npad doubles: the full x_hat this rank's fused y-phase gathers from
//   flags  2 x 16 epochs: A[q] = "rank q's partial w of epoch e is complete", B[q] = "rank q has stored its x_hat
//          block of epoch e everywhere" -- written remotely by rank q (st.release.sys), polled locally (ld.acquire.sys)
// fused_exchange_x_kernel (engine.cu): wait A -> for j in the owned x-block: w_j = sum over the P local slots in rank
// order (deterministic), x-update, x_hat_j stored into EVERY rank's xhat over P2P -> last CTA signals B.
// NVLink traffic per GPU per iteration: 8 n (P-1)/P bytes out under the SpMV pass (partial w) and the same out of the
// x-update kernel (x_hat); everything is pushed (stores), nothing is pulled.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kMaxPeers = 16;
struct PeerExchange {
    int P = 1, rank = 0, device = 0;
    size_t npad = 0;
    double *w[kMaxPeers] = {};
    double *xhat[kMaxPeers] = {};
    unsigned long long *flags[kMaxPeers] = {};   // flags[q] + 0..15: A slots, + 16..31: B slots (memory of rank q)
    unsigned *done = nullptr;                    // local: CTAs of the fused kernel that finished
    unsigned long long epoch = 0;                // exchanges issued so far (same on every rank)
    bool ipc_opened[kMaxPeers] = {};
};
// Collective calls (every rank, same order).  create returns nullptr on every rank if any rank cannot map its peers.
PeerExchange *peer_exchange_create(Collective *coll, int device, size_t npad, cudaStream_t st);
void peer_exchange_destroy(PeerExchange *px, Collective *coll, cudaStream_t st);

struct LocalGroup;   // shared state of P in-process logical ranks
LocalGroup *local_group_create(int nranks);
void local_group_destroy(LocalGroup *g);
Collective *make_local_collective(LocalGroup *g, int rank);

}  // namespace hpr
