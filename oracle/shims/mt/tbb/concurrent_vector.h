// Oracle shim (thread-pool flavour): per-worker segments, concatenated in worker order when read.
// Contention-free appends (cheaper than oneTBB's) and, with the pool's static chunking, the
// concatenation is in input-index order, so results equal the serial flavour's.
#pragma once
#include <vector>
#include "pool.h"
namespace tbb {
template <class T>
class concurrent_vector {
public:
    using iterator = typename std::vector<T>::iterator;
    using const_iterator = typename std::vector<T>::const_iterator;
    concurrent_vector() : seg_(static_cast<std::size_t>(shim::Pool::get().size())) {}
    void reserve(std::size_t n) { const std::size_t per = n / seg_.size() + 1; for (auto &s : seg_) s.reserve(per); }
    template <class... A> void emplace_back(A &&...a) { seg_[shim::worker_id()].emplace_back(std::forward<A>(a)...); dirty_ = true; }
    void push_back(const T &v) { emplace_back(v); }
    iterator begin() { flatten(); return flat_.begin(); }
    iterator end() { flatten(); return flat_.end(); }
    std::size_t size() { flatten(); return flat_.size(); }
    bool empty() { return size() == 0; }
private:
    void flatten() {
        if (!dirty_) return;
        for (auto &s : seg_) { flat_.insert(flat_.end(), std::make_move_iterator(s.begin()), std::make_move_iterator(s.end())); s.clear(); }
        dirty_ = false;
    }
    std::vector<std::vector<T>> seg_;
    std::vector<T> flat_;
    bool dirty_ = false;
};
}  // namespace tbb
