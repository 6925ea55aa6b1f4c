// Oracle shim (test infrastructure): minimal stand-in for oneTBB's blocked_range,
// which is not installed in this image. Semantics only: a half-open index range.
#pragma once
#include <cstddef>
namespace tbb {
template <class T>
class blocked_range {
public:
    blocked_range(T b, T e, std::size_t grain = 1) : b_(b), e_(e) { (void)grain; }
    T begin() const { return b_; }
    T end() const { return e_; }
    std::size_t size() const { return static_cast<std::size_t>(e_ - b_); }
    bool empty() const { return !(b_ < e_); }
private:
    T b_, e_;
};
}  // namespace tbb
