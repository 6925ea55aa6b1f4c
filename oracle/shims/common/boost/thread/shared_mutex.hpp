// Oracle shim: Boost.Thread reader/writer lock vocabulary with NULL locks.
//
// Why null: every writer on the path (insert_points, remove_points_from_far, clear) runs on the
// single caller thread while no parallel section is active (odom_run.cpp:154-185 is one thread;
// TBB fans out only inside read-only loops), so the locks never arbitrate anything. With real
// locks VoxelHashMap::remove_points_from_far takes shared_lock(map_mutex) and then
// unique_lock(map_mutex) on the same thread (voxel_hash_map.cpp:156,162) and self-deadlocks as soon
// as one voxel qualifies. Null locks let the reference's own eviction code run,
// which is what makes it usable as the oracle for that function; they also make the timed CPU
// baseline strictly faster than it would be with Boost.
#pragma once
namespace boost {
struct shared_mutex {
    void lock() {} void unlock() {} bool try_lock() { return true; }
    void lock_shared() {} void unlock_shared() {} bool try_lock_shared() { return true; }
};
struct defer_lock_t {};
constexpr defer_lock_t defer_lock{};
template <class M> struct shared_lock {
    explicit shared_lock(M &) {} shared_lock(M &, defer_lock_t) {}
    void lock() {} void unlock() {} bool try_lock() { return true; }
};
template <class M> struct unique_lock {
    explicit unique_lock(M &) {} unique_lock(M &, defer_lock_t) {}
    void lock() {} void unlock() {} bool try_lock() { return true; }
};
}  // namespace boost
