// Oracle shim (thread-pool flavour): random-access ranges fan out; forward ranges run in order.
#pragma once
#include <iterator>
#include <type_traits>
#include "pool.h"
namespace tbb {
namespace shim {
template <class It, class F>
void pfe(It first, It last, const F &f, std::random_access_iterator_tag) {
    if (first == last) return;
    Pool::get().run(static_cast<std::size_t>(last - first), [&](int, std::size_t b, std::size_t e) {
        for (std::size_t i = b; i < e; ++i) f(first[static_cast<std::ptrdiff_t>(i)]);
    });
}
template <class It, class F>
void pfe(It first, It last, const F &f, std::forward_iterator_tag) { for (; first != last; ++first) f(*first); }
}  // namespace shim
template <class It, class F>
void parallel_for_each(It first, It last, const F &f) {
    shim::pfe(first, last, f, typename std::iterator_traits<It>::iterator_category());
}
}  // namespace tbb
