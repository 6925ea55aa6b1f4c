// Oracle shim: SERIAL parallel_for. One body call over the whole range / ascending index
// loop, so every "parallel" loop of the reference runs in input-index order (deterministic).
#pragma once
#include "blocked_range.h"
namespace tbb {
template <class T, class F>
void parallel_for(const blocked_range<T> &r, const F &f) { if (!r.empty()) f(r); }
template <class I, class F>
void parallel_for(I first, I last, const F &f) { for (I i = first; i < last; ++i) f(i); }
}  // namespace tbb
