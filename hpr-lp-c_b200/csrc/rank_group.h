// rank_group.h -- P ranks of the row-partitioned mode inside ONE process: one host thread + one engine + one
// collective endpoint per rank (NCCL: one GPU per rank, ncclCommInitAll; local: all logical ranks on one device,
// see collective.h).  If a rank throws, every endpoint is aborted so that peers blocked in a collective (or in the
// stream synchronisation behind it) return instead of hanging, and the first error is re-thrown after the join.
#pragma once
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "collective.h"

namespace hpr {

enum class Transport { Nccl, Local };

class RankGroup {
   public:
    int P;
    std::vector<int> devices;
    std::vector<std::unique_ptr<Collective>> colls;
    LocalGroup *local = nullptr;

    RankGroup(Transport t, const std::vector<int> &devs) : P((int)devs.size()), devices(devs) {
        if (t == Transport::Nccl) {
            std::vector<NcclComm> comms(P, nullptr);
            const int rc = nccl().CommInitAll(comms.data(), P, devs.data());
            if (rc != 0) throw std::runtime_error(std::string("ncclCommInitAll failed: ") + nccl().GetErrorString(rc));
            for (int p = 0; p < P; ++p) colls.emplace_back(make_nccl_collective(comms[p], P, p));
        } else {
            local = local_group_create(P);
            for (int p = 0; p < P; ++p) colls.emplace_back(make_local_collective(local, p));
        }
    }
    ~RankGroup() {
        colls.clear();
        if (local) local_group_destroy(local);
    }
    RankGroup(const RankGroup &) = delete;
    RankGroup &operator=(const RankGroup &) = delete;

    template <class F>
    void run(F &&per_rank) {   // per_rank(p, device, collective)
        std::vector<std::string> errors(P);
        std::vector<std::thread> workers;
        for (int p = 0; p < P; ++p) {
            workers.emplace_back([&, p]() {
                try {
                    per_rank(p, devices[p], colls[p].get());
                } catch (const std::exception &e) {
                    errors[p] = e.what();
                    for (auto &c : colls) c->abort();
                } catch (...) {
                    errors[p] = "unknown exception";
                    for (auto &c : colls) c->abort();
                }
            });
        }
        for (auto &w : workers) w.join();
        for (int p = 0; p < P; ++p)   // the first failure is the cause; later ones are usually "peer rank failed"
            if (!errors[p].empty() && errors[p].find("peer rank failed") == std::string::npos)
                throw std::runtime_error("rank " + std::to_string(p) + ": " + errors[p]);
        for (int p = 0; p < P; ++p)
            if (!errors[p].empty()) throw std::runtime_error("rank " + std::to_string(p) + ": " + errors[p]);
    }
};

}  // namespace hpr
