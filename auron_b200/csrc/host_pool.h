// host_pool.h -- process-wide host-side helpers shared by the scan and the shuffle writer: a persistent worker pool
// (parallel_for) and a pool of pinned staging buffers.
#pragma once
#include <sched.h>
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "common.h"

namespace auron {

// CPUs this process may run on (sched_getaffinity): with one process per GPU bound to the GPU's NUMA node that is the node's
// share of the box, not all of it.  Sizing the pools by hardware_concurrency() put 8 x 64 workers on 128 CPUs at 8 GPUs.
inline unsigned usable_cpus() {
    cpu_set_t set;
    CPU_ZERO(&set);
    if (sched_getaffinity(0, sizeof(set), &set) == 0) {
        const int n = CPU_COUNT(&set);
        if (n > 0) return (unsigned)n;
    }
    const unsigned h = std::thread::hardware_concurrency();
    return h ? h : 1u;
}

// Persistent host worker pool (process-wide).  Spawning 32 std::threads per parallel_for cost ~0.6 ms per scan batch,
// more than the page-header parsing they were spawned for.
class WorkerPool {
  public:
    struct Job {
        std::function<void()> work;
        int outstanding = 0;   // tickets handed to the pool and not yet finished (guarded by pool mutex)
    };
    static WorkerPool& get() {
        static WorkerPool* pool = new WorkerPool();   // leaked on purpose: workers may outlive static destruction
        return *pool;
    }
    // run job.work() on up to `extra` pool workers in addition to the caller; returns when all of them are done
    void run(Job& job, unsigned extra) {
        extra = std::min<unsigned>(extra, (unsigned)workers_.size());
        {
            std::lock_guard<std::mutex> l(mu_);
            job.outstanding = (int)extra;
            for (unsigned i = 0; i < extra; i++) tickets_.push_back(&job);
        }
        if (extra == 1) cv_.notify_one();
        else if (extra > 1) cv_.notify_all();
        job.work();
        std::unique_lock<std::mutex> l(mu_);
        for (auto it = tickets_.begin(); it != tickets_.end();) {   // tickets nobody picked up yet are not needed any more
            if (*it == &job) {
                it = tickets_.erase(it);
                job.outstanding--;
            } else ++it;
        }
        done_cv_.wait(l, [&] { return job.outstanding == 0; });
    }

  private:
    WorkerPool() {
        unsigned n = std::max(1u, std::min(64u, usable_cpus()));
        for (unsigned i = 0; i < n; i++) {
            workers_.emplace_back([this] { loop(); });
            workers_.back().detach();
        }
    }
    void loop() {
        for (;;) {
            Job* j;
            {
                std::unique_lock<std::mutex> l(mu_);
                cv_.wait(l, [&] { return !tickets_.empty(); });
                j = tickets_.front();
                tickets_.pop_front();
            }
            j->work();
            std::lock_guard<std::mutex> l(mu_);
            if (--j->outstanding == 0) done_cv_.notify_all();
        }
    }
    std::mutex mu_;
    std::condition_variable cv_, done_cv_;
    std::deque<Job*> tickets_;
    std::vector<std::thread> workers_;
};

// run fn(i) for i in [0, n) on up to `threads` host threads; the first exception is rethrown
template <typename F>
inline void parallel_for(size_t n, unsigned threads, F fn) {
    if (n == 0) return;
    threads = (unsigned)std::max<size_t>(1, std::min<size_t>(threads, n));
    if (threads == 1) {
        for (size_t i = 0; i < n; i++) fn(i);
        return;
    }
    std::atomic<size_t> next{0};
    std::mutex mu;
    std::string err;
    WorkerPool::Job job;
    job.work = [&]() {
        for (;;) {
            size_t i = next.fetch_add(1);
            if (i >= n) return;
            try {
                fn(i);
            } catch (const std::exception& e) {
                std::lock_guard<std::mutex> l(mu);
                if (err.empty()) err = e.what();
            } catch (...) {
                std::lock_guard<std::mutex> l(mu);
                if (err.empty()) err = "unknown failure in a scan worker";
            }
        }
    };
    WorkerPool::get().run(job, threads - 1);
    if (!err.empty()) fail(err);
}


// pinned staging comes from a process-wide pool: cudaHostAlloc costs ~0.4 s per GB, far more than the copy it feeds
struct PinnedPool {
    std::mutex mu;
    std::vector<std::pair<void*, size_t>> free_list;
    size_t total = 0;   // bytes handed out + bytes in the free list
    // cudaHostAlloc / cudaFreeHost cost ~0.3 ms per MB: blocks are kept and reused (best fit); idle blocks are only released when
    // the pool would otherwise exceed its cap (AURON_PINNED_POOL_BYTES, default 12 GB), smallest first
    static size_t cap_bytes() {
        static const size_t c = getenv("AURON_PINNED_POOL_BYTES") ? (size_t)atoll(getenv("AURON_PINNED_POOL_BYTES")) : (size_t)12 << 30;
        return c;
    }
    void* get(size_t n, size_t* cap) {
        std::lock_guard<std::mutex> l(mu);
        size_t best = SIZE_MAX;
        for (size_t i = 0; i < free_list.size(); i++)
            if (free_list[i].second >= n && (best == SIZE_MAX || free_list[i].second < free_list[best].second)) best = i;
        if (best != SIZE_MAX) {
            auto e = free_list[best];
            free_list.erase(free_list.begin() + best);
            *cap = e.second;
            return e.first;
        }
        const size_t c = std::max<size_t>(n + n / 8, 64 << 20);
        while (total + c > cap_bytes() && !free_list.empty()) {   // make room: idle blocks that are too small for this request
            size_t smallest = 0;
            for (size_t i = 1; i < free_list.size(); i++)
                if (free_list[i].second < free_list[smallest].second) smallest = i;
            cudaFreeHost(free_list[smallest].first);
            total -= free_list[smallest].second;
            free_list.erase(free_list.begin() + smallest);
        }
        void* p = nullptr;
        CUDA_OK(cudaHostAlloc(&p, c, cudaHostAllocDefault));
        total += c;
        *cap = c;
        return p;
    }
    void put(void* p, size_t cap) {
        std::lock_guard<std::mutex> l(mu);
        free_list.emplace_back(p, cap);
    }
};
inline PinnedPool& pinned_pool() {
    static PinnedPool pool;
    return pool;
}

}  // namespace auron
