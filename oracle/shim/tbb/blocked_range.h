// ORACLE — TEST INFRASTRUCTURE ONLY.  NOT oneTBB.  Stand-in for tbb::blocked_range (oneTBB v2021.8.0 in the reference,
// cpp/sage_icp/3rdparty/tbb/tbb.cmake:31; not in this image): a half-open range [begin, end) that is never split.
#pragma once
#include <cstddef>

namespace tbb {
template <class Value>
class blocked_range {
public:
    blocked_range(Value b, Value e) : b_(b), e_(e) {}
    Value begin() const { return b_; }
    Value end() const { return e_; }
    std::size_t size() const { return (std::size_t)(e_ - b_); }
    bool empty() const { return !(b_ < e_); }

private:
    Value b_, e_;
};
}  // namespace tbb
