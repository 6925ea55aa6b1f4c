// Oracle shim: parallel_sort = std::sort.
#pragma once
#include <algorithm>
namespace tbb {
template <class It> void parallel_sort(It b, It e) { std::sort(b, e); }
template <class It, class C> void parallel_sort(It b, It e, C c) { std::sort(b, e, c); }
}  // namespace tbb
