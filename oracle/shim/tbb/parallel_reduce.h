// ORACLE — TEST INFRASTRUCTURE ONLY.  NOT oneTBB.
//
// tbb::parallel_reduce(range, identity, body, join) over SAGE_REF_THREADS contiguous chunks (default 1 = the calling thread
// only), partial results joined left to right — a legal TBB schedule: for GetCorrespondences the pairs come out in query order
// whatever the thread count, for AlignClouds only the summation tree changes.  With one thread the reference's results are those
// of its sequential order, which is what the parity tests use; bench.py's reference arm sets SAGE_REF_THREADS to the host's
// cores.
#pragma once
#include <cstdlib>
#include <thread>
#include <vector>

#include "blocked_range.h"

namespace tbb {

inline int shim_threads() {
    const char *e = std::getenv("SAGE_REF_THREADS");
    const int n = e ? std::atoi(e) : 1;
    return n < 1 ? 1 : n;
}

template <class Range, class Value, class Body, class Join>
Value parallel_reduce(const Range &range, const Value &identity, const Body &body, const Join &join) {
    const int T = shim_threads();
    const std::size_t n = range.size();
    if (T <= 1 || n < (std::size_t)(4 * T)) return body(range, identity);
    std::vector<Value> parts((std::size_t)T, identity);
    std::vector<std::thread> workers;
    for (int t = 0; t < T; ++t)
        workers.emplace_back([&, t] {
            const Range chunk(range.begin() + (n * (std::size_t)t) / T, range.begin() + (n * (std::size_t)(t + 1)) / T);
            parts[(std::size_t)t] = body(chunk, identity);
        });
    for (auto &w : workers) w.join();
    Value acc = parts[0];
    for (int t = 1; t < T; ++t) acc = join(acc, parts[(std::size_t)t]);
    return acc;
}

}  // namespace tbb
