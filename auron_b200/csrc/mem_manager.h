// mem_manager.h -- the HBM tier's analogue of the reference's memory manager (native-engine/auron-memmgr/src/lib.rs:201-423).
// The reference hands every spillable operator (aggregate table, sorter, shuffle repartitioner) a share of one executor-wide budget:
// a consumer reports its usage after every change and is told to spill when the total is over the budget and it holds at least its
// fair share.  Here the budget is a fraction of the device memory (40 %: inputs, scratch and the stream-ordered pools keep the rest;
// AURON_HBM_BUDGET_BYTES overrides), shared by all tasks of the process on that device, and "spill" moves the consumer's state to
// pinned host memory (AggExec buckets, SortExec runs).  Consumers do not wait for each other (every task runs on its own thread and
// stream): one that is over its share spills itself, one that is under it goes on and the holders above the share spill at
// their next update.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>
#include <map>
#include <mutex>
#include <string>

#include "common.h"

namespace auron {

class MemManager {
   public:
    static MemManager& of(int device) {
        static std::mutex mu;
        static std::map<int, MemManager*> all;
        std::lock_guard<std::mutex> g(mu);
        auto it = all.find(device);
        if (it == all.end()) it = all.emplace(device, new MemManager(device)).first;
        return *it->second;
    }
    int add(const std::string& name) {
        std::lock_guard<std::mutex> g(mu_);
        const int id = next_id_++;
        used_[id] = {name, 0};
        return id;
    }
    void remove(int id) {
        std::lock_guard<std::mutex> g(mu_);
        used_.erase(id);
    }
    // consumer `id` now holds `bytes` of spillable state; true = spill it now
    bool update(int id, int64_t bytes) {
        std::lock_guard<std::mutex> g(mu_);
        auto it = used_.find(id);
        if (it == used_.end()) return false;
        it->second.second = bytes;
        int64_t total = 0, holders = 0;
        for (auto& kv : used_) {
            total += kv.second.second;
            holders += kv.second.second > 0;
        }
        if (total <= budget_ || bytes <= 0) return false;
        return bytes >= budget_ / (2 * std::max<int64_t>(holders, 1));   // at least half a fair share: worth spilling
    }
    int64_t budget() const { return budget_; }
    int64_t set_budget(int64_t bytes) {   // <= 0: back to the default
        std::lock_guard<std::mutex> g(mu_);
        budget_ = bytes > 0 ? bytes : default_budget_;
        return budget_;
    }
    int64_t total_used() {
        std::lock_guard<std::mutex> g(mu_);
        int64_t total = 0;
        for (auto& kv : used_) total += kv.second.second;
        return total;
    }

   private:
    explicit MemManager(int device) {
        if (const char* e = getenv("AURON_HBM_BUDGET_BYTES")) budget_ = atoll(e);
        if (budget_ <= 0) {
            // asked once per process and device: cudaMemGetInfo takes the driver's context lock, which a scan's copy thread would wait behind
            int cur = 0;
            cudaGetDevice(&cur);
            cudaSetDevice(device);
            size_t free_b = 0, total_b = 0;
            CUDA_OK(cudaMemGetInfo(&free_b, &total_b));
            cudaSetDevice(cur);
            budget_ = (int64_t)(total_b / 10 * 4);
        }
        default_budget_ = budget_;
    }
    std::mutex mu_;
    std::map<int, std::pair<std::string, int64_t>> used_;
    int next_id_ = 1;
    int64_t budget_ = 0, default_budget_ = 0;
};

}  // namespace auron
