// ORACLE — TEST INFRASTRUCTURE ONLY.  NOT oneTBB.  tbb::parallel_for(first, last, f) over SAGE_REF_THREADS contiguous chunks
// (default 1: a plain loop on the calling thread) on the worker pool of parallel_reduce.h.
#pragma once
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
    shim_pool::instance().run(T, [&](int t) {
        const Index b = first + (Index)((n * (std::size_t)t) / T), e = first + (Index)((n * (std::size_t)(t + 1)) / T);
        for (Index i = b; i < e; ++i) f(i);
    });
}
}  // namespace tbb
