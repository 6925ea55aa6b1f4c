// Oracle shim: SERIAL parallel_reduce = body(whole range, identity): a single
// left-to-right accumulation. This defines the oracle's rounding of H and g.
#pragma once
#include "blocked_range.h"
namespace tbb {
template <class R, class V, class B, class J>
V parallel_reduce(const R &range, const V &identity, const B &body, const J &join) {
    (void)join;
    return body(range, identity);
}
}  // namespace tbb
