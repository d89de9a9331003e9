// collective.cu -- transports of the row-partitioned exchange steps (see collective.h).
#include "collective.h"

#include <atomic>
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

LocalGroup *local_group_create(int nranks) {
    if (nranks < 1 || nranks > kMaxLocalRanks) throw std::runtime_error("local collective: 1..16 ranks");
    LocalGroup *g = new LocalGroup;
    g->P = nranks;
    return g;
}
void local_group_destroy(LocalGroup *g) { delete g; }
Collective *make_local_collective(LocalGroup *g, int rank) { return new LocalCollective(g, rank); }

}  // namespace hpr
