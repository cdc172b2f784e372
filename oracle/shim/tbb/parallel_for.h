// ORACLE — TEST INFRASTRUCTURE ONLY.  NOT oneTBB.  tbb::parallel_for(first, last, f) as a plain loop on the calling thread.
#pragma once
#include "blocked_range.h"

namespace tbb {
template <class Index, class F>
void parallel_for(Index first, Index last, const F &f) {
    for (Index i = first; i < last; ++i) f(i);
}
}  // namespace tbb
