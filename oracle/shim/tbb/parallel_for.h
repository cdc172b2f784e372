// ORACLE — TEST INFRASTRUCTURE ONLY.  NOT oneTBB.  tbb::parallel_for(first, last, f) over SAGE_REF_THREADS contiguous chunks
// (default 1: a plain loop on the calling thread); see parallel_reduce.h.
#pragma once
#include <thread>
#include <vector>

#include "parallel_reduce.h"

namespace tbb {
template <class Index, class F>
void parallel_for(Index first, Index last, const F &f) {
    const int T = shim_threads();
    const std::size_t n = last > first ? (std::size_t)(last - first) : 0;
    if (T <= 1 || n < (std::size_t)(4 * T)) {
        for (Index i = first; i < last; ++i) f(i);
        return;
    }
    std::vector<std::thread> workers;
    for (int t = 0; t < T; ++t)
        workers.emplace_back([&, t] {
            const Index b = first + (Index)((n * (std::size_t)t) / T), e = first + (Index)((n * (std::size_t)(t + 1)) / T);
            for (Index i = b; i < e; ++i) f(i);
        });
    for (auto &w : workers) w.join();
}
}  // namespace tbb
