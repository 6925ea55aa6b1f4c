// Oracle shim (thread-pool flavour): one partial per worker over a contiguous sub-range, joined
// in worker order (deterministic for a fixed thread count; rounding differs from the serial sum).
#pragma once
#include <vector>
#include "blocked_range.h"
#include "pool.h"
namespace tbb {
template <class T, class V, class B, class J>
V parallel_reduce(const blocked_range<T> &range, const V &identity, const B &body, const J &join) {
    if (range.empty()) return identity;
    const int n = shim::Pool::get().size();
    std::vector<V> part(static_cast<std::size_t>(n), identity);
    std::vector<char> used(static_cast<std::size_t>(n), 0);
    shim::Pool::get().run(range.size(), [&](int w, std::size_t b, std::size_t e) {
        part[w] = body(blocked_range<T>(range.begin() + static_cast<T>(b), range.begin() + static_cast<T>(e)), identity);
        used[w] = 1;
    });
    V acc = identity;
    bool first = true;
    for (int w = 0; w < n; ++w) {
        if (!used[w]) continue;
        if (first) { acc = part[w]; first = false; } else acc = join(acc, part[w]);
    }
    return acc;
}
}  // namespace tbb
