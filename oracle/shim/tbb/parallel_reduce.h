// ORACLE — TEST INFRASTRUCTURE ONLY.  NOT oneTBB.  tbb::parallel_reduce(range, identity, body, join) run as ONE chunk on the
// calling thread: a legal TBB schedule (the one a single-threaded arena produces), so the reference's results are those of
// its sequential order — for GetCorrespondences the pairs come out in query order, for AlignClouds the sums are left to right.
#pragma once
#include "blocked_range.h"

namespace tbb {
template <class Range, class Value, class Body, class Join>
Value parallel_reduce(const Range &range, const Value &identity, const Body &body, const Join &) {
    return body(range, identity);
}
}  // namespace tbb
