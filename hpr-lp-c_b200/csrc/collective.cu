// collective.cu -- transports of the row-partitioned exchange steps (see collective.h).
#include "collective.h"

#include <unistd.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <condition_variable>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

namespace hpr {

namespace {

void check_cuda(cudaError_t e, const char *what) {
    if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}

// ---------------------------------------------------------------------------------------------
// NCCL
// ---------------------------------------------------------------------------------------------
class NcclCollective : public Collective {
   public:
    std::atomic<NcclComm> comm;   // owned; abort() may come from another rank's host thread
    NcclCollective(NcclComm c, int P, int r) : comm(c) { nranks = P; rank = r; }
    ~NcclCollective() override {
        NcclComm c = comm.exchange(nullptr);
        if (c) nccl().CommDestroy(c);
    }
    NcclComm live() {
        NcclComm c = comm.load();
        if (!c) throw std::runtime_error("NCCL communicator was aborted (a peer rank failed)");
        return c;
    }
    void check(int rc, const char *what) {
        if (rc != 0) throw std::runtime_error(std::string(what) + " failed: " + nccl().GetErrorString(rc));
    }
    void all_reduce(double *buf, size_t count, bool max_op, cudaStream_t st) override {
        check(nccl().AllReduce(buf, buf, count, kNcclFloat64, max_op ? kNcclMax : kNcclSum, live(), st), "ncclAllReduce");
    }
    void reduce_scatter_inplace(double *buf, size_t block, cudaStream_t st) override {
        check(nccl().ReduceScatter(buf, buf + (size_t)rank * block, block, kNcclFloat64, kNcclSum, live(), st), "ncclReduceScatter");
    }
    void all_gather_inplace(double *buf, size_t block, cudaStream_t st) override {
        check(nccl().AllGather(buf + (size_t)rank * block, buf, block, kNcclFloat64, live(), st), "ncclAllGather");
    }
    void abort() override {
        NcclComm c = comm.exchange(nullptr);
        if (c) nccl().CommAbort(c);   // unblocks a peer thread stuck in (or synchronising on) a collective of this communicator
    }
    const char *name() const override { return "nccl"; }
};

// ---------------------------------------------------------------------------------------------
// local: directly addressable buffers, host barriers, rank-ordered kernels
// ---------------------------------------------------------------------------------------------
constexpr int kMaxLocalRanks = 16;
struct PtrTable { const double *p[kMaxLocalRanks]; };

__global__ void local_reduce_block_kernel(PtrTable t, int P, size_t offset, size_t count, double *out, bool max_op) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
        double v = t.p[0][offset + i];
        for (int q = 1; q < P; ++q) {
            const double w = t.p[q][offset + i];
            v = max_op ? fmax(v, w) : v + w;   // rank order: deterministic
        }
        out[i] = v;
    }
}
__global__ void local_gather_kernel(PtrTable t, int P, int self, size_t block, double *buf) {
    for (int q = 0; q < P; ++q) {
        if (q == self) continue;
        const double *src = t.p[q] + (size_t)q * block;
        double *dst = buf + (size_t)q * block;
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < block; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
    }
}

}  // namespace

struct LocalGroup {
    int P = 1;
    std::mutex mu;
    std::condition_variable cv;
    int arrived = 0;
    unsigned long long gen = 0;
    bool failed = false;
    PtrTable table{};
    void barrier() {
        std::unique_lock<std::mutex> lk(mu);
        if (failed) throw std::runtime_error("local collective: a peer rank failed");
        const unsigned long long g = gen;
        if (++arrived == P) { arrived = 0; ++gen; cv.notify_all(); return; }
        cv.wait(lk, [&] { return gen != g || failed; });
        if (failed && gen == g) throw std::runtime_error("local collective: a peer rank failed");
    }
    void poison() {
        std::lock_guard<std::mutex> lk(mu);
        failed = true;
        cv.notify_all();
    }
};

namespace {

class LocalCollective : public Collective {
   public:
    LocalGroup *g;
    double *scratch = nullptr;
    size_t scratch_count = 0;
    LocalCollective(LocalGroup *grp, int r) : g(grp) { nranks = grp->P; rank = r; }
    ~LocalCollective() override { if (scratch) cudaFree(scratch); }
    static int grid_for(size_t count) { return (int)std::max<size_t>(1, std::min<size_t>((count + 255) / 256, 148 * 8)); }
    // publish my buffer once everything queued on my stream has finished, wait for every peer to do the same
    PtrTable enter(const double *buf, cudaStream_t st) {
        check_cuda(cudaStreamSynchronize(st), "local collective sync");
        { std::lock_guard<std::mutex> lk(g->mu); g->table.p[rank] = buf; }
        g->barrier();
        PtrTable t;
        { std::lock_guard<std::mutex> lk(g->mu); t = g->table; }
        return t;
    }
    // my reads of the peers' buffers are done; nobody may touch its buffer again before all are
    void leave(cudaStream_t st) {
        check_cuda(cudaGetLastError(), "local collective kernel");
        check_cuda(cudaStreamSynchronize(st), "local collective sync");
        g->barrier();
    }
    void all_reduce(double *buf, size_t count, bool max_op, cudaStream_t st) override {
        if (count > scratch_count) {
            if (scratch) cudaFree(scratch);
            check_cuda(cudaMalloc(&scratch, sizeof(double) * count), "local collective scratch");
            scratch_count = count;
        }
        const PtrTable t = enter(buf, st);
        local_reduce_block_kernel<<<grid_for(count), 256, 0, st>>>(t, nranks, 0, count, scratch, max_op);
        leave(st);
        check_cuda(cudaMemcpyAsync(buf, scratch, sizeof(double) * count, cudaMemcpyDeviceToDevice, st), "local collective copy");
    }
    void reduce_scatter_inplace(double *buf, size_t block, cudaStream_t st) override {
        const PtrTable t = enter(buf, st);
        // rank r reads block r of every rank and writes block r of its own buffer, which no peer reads
        local_reduce_block_kernel<<<grid_for(block), 256, 0, st>>>(t, nranks, (size_t)rank * block, block, buf + (size_t)rank * block, false);
        leave(st);
    }
    void all_gather_inplace(double *buf, size_t block, cudaStream_t st) override {
        const PtrTable t = enter(buf, st);
        local_gather_kernel<<<grid_for(block), 256, 0, st>>>(t, nranks, rank, block, buf);
        leave(st);
    }
    void abort() override { g->poison(); }
    const char *name() const override { return "local"; }
};

}  // namespace

Collective *make_nccl_collective(NcclComm comm, int nranks, int rank) { return new NcclCollective(comm, nranks, rank); }

// ---------------------------------------------------------------------------------------------------------------
// peer-memory exchange bootstrap
// ---------------------------------------------------------------------------------------------------------------
namespace {
struct PeerRecord {   // what a rank tells the others about its exchange buffers (padded to 32 doubles)
    long long pid;
    int device, ok;
    void *w, *xhat, *flags;
    cudaIpcMemHandle_t hw, hx, hf;
};
static_assert(sizeof(PeerRecord) <= 32 * sizeof(double), "PeerRecord must fit its all-gather slot");
constexpr size_t kRecDoubles = 32;

// every rank contributes ok (1/0); returns true iff all ranks said 1
bool all_ranks_ok(Collective *coll, bool ok, double *dbuf, cudaStream_t st) {
    double v = ok ? 0.0 : 1.0;
    check_cuda(cudaMemcpyAsync(dbuf, &v, sizeof(double), cudaMemcpyHostToDevice, st), "peer exchange vote");
    coll->all_reduce(dbuf, 1, false, st);
    check_cuda(cudaMemcpyAsync(&v, dbuf, sizeof(double), cudaMemcpyDeviceToHost, st), "peer exchange vote");
    check_cuda(cudaStreamSynchronize(st), "peer exchange vote");
    return v == 0.0;
}
}  // namespace

PeerExchange *peer_exchange_create(Collective *coll, int device, size_t npad, cudaStream_t st) {
    const int P = coll->nranks, rank = coll->rank;
    if (P < 2 || P > kMaxPeers || std::string(coll->name()) != "nccl") return nullptr;
    if (const char *e = getenv("HPRLP_EXCHANGE"))
        if (std::string(e) == "nccl") return nullptr;   // same value on every rank (environment of one job)
    std::unique_ptr<PeerExchange> px(new PeerExchange);
    px->P = P; px->rank = rank; px->device = device; px->npad = npad;
    double *dbuf = nullptr;   // bootstrap all-gather buffer
    auto release_all = [&]() {   // local teardown of whatever exists so far (error paths; nobody else uses the buffers yet)
        for (int q = 0; q < P; ++q)
            if (px->ipc_opened[q]) { cudaIpcCloseMemHandle(px->w[q]); cudaIpcCloseMemHandle(px->xhat[q]); cudaIpcCloseMemHandle(px->flags[q]); }
        cudaFree(px->w[rank]); cudaFree(px->xhat[rank]); cudaFree(px->flags[rank]); cudaFree(px->done);
        cudaFree(dbuf);
        cudaGetLastError();
    };
    try {
    check_cuda(cudaMalloc(&dbuf, sizeof(double) * kRecDoubles * P), "peer exchange bootstrap");
    PeerRecord mine{};
    mine.pid = (long long)getpid();
    mine.device = device;
    bool ok = true;
    ok = ok && cudaMalloc(&px->w[rank], sizeof(double) * npad) == cudaSuccess;
    ok = ok && cudaMalloc(&px->xhat[rank], sizeof(double) * npad) == cudaSuccess;
    ok = ok && cudaMalloc(&px->flags[rank], sizeof(unsigned long long) * 2 * kMaxPeers) == cudaSuccess;
    ok = ok && cudaMalloc(&px->done, sizeof(unsigned)) == cudaSuccess;
    if (ok) {
        cudaMemsetAsync(px->w[rank], 0, sizeof(double) * npad, st);
        cudaMemsetAsync(px->xhat[rank], 0, sizeof(double) * npad, st);
        cudaMemsetAsync(px->flags[rank], 0, sizeof(unsigned long long) * 2 * kMaxPeers, st);
        cudaMemsetAsync(px->done, 0, sizeof(unsigned), st);
        mine.w = px->w[rank]; mine.xhat = px->xhat[rank]; mine.flags = px->flags[rank];
        ok = ok && cudaIpcGetMemHandle(&mine.hw, px->w[rank]) == cudaSuccess;
        ok = ok && cudaIpcGetMemHandle(&mine.hx, px->xhat[rank]) == cudaSuccess;
        ok = ok && cudaIpcGetMemHandle(&mine.hf, px->flags[rank]) == cudaSuccess;
    }
    cudaGetLastError();
    mine.ok = ok ? 1 : 0;
    std::vector<double> host(kRecDoubles * P, 0.0);
    std::memcpy(host.data() + kRecDoubles * rank, &mine, sizeof(mine));
    check_cuda(cudaMemcpyAsync(dbuf + kRecDoubles * rank, host.data() + kRecDoubles * rank, sizeof(double) * kRecDoubles, cudaMemcpyHostToDevice, st), "peer exchange bootstrap");
    coll->all_gather_inplace(dbuf, kRecDoubles, st);
    check_cuda(cudaMemcpyAsync(host.data(), dbuf, sizeof(double) * kRecDoubles * P, cudaMemcpyDeviceToHost, st), "peer exchange bootstrap");
    check_cuda(cudaStreamSynchronize(st), "peer exchange bootstrap");
    for (int q = 0; q < P && ok; ++q) {
        PeerRecord r;
        std::memcpy(&r, host.data() + kRecDoubles * q, sizeof(r));
        if (!r.ok) { ok = false; break; }
        if (q == rank) continue;
        if (r.pid == mine.pid) {   // same process (one host thread per GPU): raw pointers + peer access
            int can = 0;
            if (r.device != device) {
                cudaDeviceCanAccessPeer(&can, device, r.device);
                if (!can) { ok = false; break; }
                const cudaError_t e = cudaDeviceEnablePeerAccess(r.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { ok = false; break; }
                cudaGetLastError();
            }
            px->w[q] = static_cast<double *>(r.w); px->xhat[q] = static_cast<double *>(r.xhat);
            px->flags[q] = static_cast<unsigned long long *>(r.flags);
        } else {                   // another process: CUDA IPC (enables peer access lazily)
            void *a = nullptr, *b = nullptr, *c = nullptr;
            if (cudaIpcOpenMemHandle(&a, r.hw, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
                cudaIpcOpenMemHandle(&b, r.hx, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
                cudaIpcOpenMemHandle(&c, r.hf, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                if (a) cudaIpcCloseMemHandle(a);
                if (b) cudaIpcCloseMemHandle(b);
                cudaGetLastError();
                ok = false;
                break;
            }
            px->w[q] = static_cast<double *>(a); px->xhat[q] = static_cast<double *>(b);
            px->flags[q] = static_cast<unsigned long long *>(c);
            px->ipc_opened[q] = true;
        }
    }
    const bool all_ok = all_ranks_ok(coll, ok, dbuf, st);
    if (!all_ok) {
        if (rank == 0) std::fprintf(stderr, "[hprlp] peer-memory exchange unavailable (no P2P / IPC mapping): using NCCL reduce-scatter + all-gather\n");
        release_all();   // nobody uses the buffers: plain local teardown (the vote above was the barrier)
        return nullptr;
    }
    cudaFree(dbuf);
    return px.release();
    } catch (...) {   // a CUDA call or a collective failed (e.g. a peer rank died and the communicator was aborted)
        release_all();
        throw;
    }
}

void peer_exchange_destroy(PeerExchange *px, Collective *coll, cudaStream_t st) {
    if (!px) return;
    // nobody may unmap or free while a peer can still touch the buffers: agree that everyone is done first
    try {
        double *dbuf = nullptr;
        if (cudaMalloc(&dbuf, sizeof(double)) == cudaSuccess) {
            all_ranks_ok(coll, true, dbuf, st);
            cudaFree(dbuf);
        }
    } catch (...) {
        // a peer failed: its communicator was aborted; fall through and release what is ours
    }
    for (int q = 0; q < px->P; ++q)
        if (px->ipc_opened[q]) { cudaIpcCloseMemHandle(px->w[q]); cudaIpcCloseMemHandle(px->xhat[q]); cudaIpcCloseMemHandle(px->flags[q]); }
    cudaFree(px->w[px->rank]); cudaFree(px->xhat[px->rank]); cudaFree(px->flags[px->rank]); cudaFree(px->done);
    cudaGetLastError();
    delete px;
}

LocalGroup *local_group_create(int nranks) {
    if (nranks < 1 || nranks > kMaxLocalRanks) throw std::runtime_error("local collective: 1..16 ranks");
    LocalGroup *g = new LocalGroup;
    g->P = nranks;
    return g;
}
void local_group_destroy(LocalGroup *g) { delete g; }
Collective *make_local_collective(LocalGroup *g, int rank) { return new LocalCollective(g, rank); }

}  // namespace hpr
