// Oracle shim: SERIAL parallel_for_each (in-order).
#pragma once
namespace tbb {
template <class It, class F>
void parallel_for_each(It first, It last, const F &f) { for (; first != last; ++first) f(*first); }
}  // namespace tbb
