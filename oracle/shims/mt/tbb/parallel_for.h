// Oracle shim (thread-pool flavour). The index overload fans out over the pool; the
// blocked_range overload stays one body call: its two call sites are the chunk-key loop of
// insert_points (voxel_hash_map.cpp:29-45, negligible work before a serial section) and
// remove_points_from_far (:152-170), whose body erases from the shared map and cannot be split.
#pragma once
#include "blocked_range.h"
#include "pool.h"
namespace tbb {
template <class T, class F>
void parallel_for(const blocked_range<T> &r, const F &f) { if (!r.empty()) f(r); }
template <class I, class F>
void parallel_for(I first, I last, const F &f) {
    if (!(first < last)) return;
    shim::Pool::get().run(static_cast<std::size_t>(last - first), [&](int, std::size_t b, std::size_t e) {
        for (std::size_t i = b; i < e; ++i) f(static_cast<I>(first + static_cast<I>(i)));
    });
}
}  // namespace tbb
