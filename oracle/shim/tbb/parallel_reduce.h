// ORACLE — TEST INFRASTRUCTURE ONLY.  NOT oneTBB.
//
// tbb::parallel_reduce(range, identity, body, join) over SAGE_REF_THREADS contiguous chunks (default 1 = the calling thread
// only), partial results joined left to right — a legal TBB schedule: for GetCorrespondences the pairs come out in query order
// whatever the thread count, for AlignClouds only the summation tree changes.  With one thread the reference's results are those
// of its sequential order, which is what the parity tests use; bench.py's reference arm sets SAGE_REF_THREADS to the host's
// cores.  Chunks run on a small persistent worker pool (like TBB's arena, threads are not created per call).
#pragma once
#include <condition_variable>
#include <cstdlib>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "blocked_range.h"

namespace tbb {

inline int shim_threads() {
    const char *e = std::getenv("SAGE_REF_THREADS");
    const int n = e ? std::atoi(e) : 1;
    return n < 1 ? 1 : (n > 256 ? 256 : n);
}

// run f(0..T-1): f(0) on the caller, the rest on pooled workers; returns when all are done
class shim_pool {
public:
    static shim_pool &instance() {
        static shim_pool p;
        return p;
    }
    void run(int T, const std::function<void(int)> &f) {
        std::lock_guard<std::mutex> serial(call_);  // one parallel region at a time (the reference never nests them)
        {
            std::unique_lock<std::mutex> lk(m_);
            while ((int)workers_.size() < T - 1) {
                const int id = (int)workers_.size() + 1;
                workers_.emplace_back([this, id] { loop(id); });
            }
            job_ = &f, width_ = T, pending_ = T - 1, ++generation_;
        }
        cv_.notify_all();
        f(0);
        std::unique_lock<std::mutex> lk(m_);
        done_.wait(lk, [this] { return pending_ == 0; });
        job_ = nullptr;
    }
    ~shim_pool() {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto &w : workers_) w.join();
    }

private:
    void loop(int id) {
        long seen = 0;
        while (true) {
            const std::function<void(int)> *job = nullptr;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return stop_ || generation_ != seen; });
                if (stop_) return;
                seen = generation_;
                if (id < width_) job = job_;
            }
            if (job) {
                (*job)(id);
                std::lock_guard<std::mutex> lk(m_);
                if (--pending_ == 0) done_.notify_one();
            }
        }
    }
    std::mutex m_, call_;
    std::condition_variable cv_, done_;
    std::vector<std::thread> workers_;
    const std::function<void(int)> *job_ = nullptr;
    int width_ = 0, pending_ = 0;
    long generation_ = 0;
    bool stop_ = false;
};

template <class Range, class Value, class Body, class Join>
Value parallel_reduce(const Range &range, const Value &identity, const Body &body, const Join &join) {
    const int T = shim_threads();
    const std::size_t n = range.size();
    if (T <= 1 || n < (std::size_t)(4 * T)) return body(range, identity);
    std::vector<Value> parts((std::size_t)T, identity);
    shim_pool::instance().run(T, [&](int t) {
        const Range chunk(range.begin() + (n * (std::size_t)t) / T, range.begin() + (n * (std::size_t)(t + 1)) / T);
        parts[(std::size_t)t] = body(chunk, identity);
    });
    Value acc = parts[0];
    for (int t = 1; t < T; ++t) acc = join(acc, parts[(std::size_t)t]);
    return acc;
}

}  // namespace tbb
