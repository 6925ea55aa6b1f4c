// Drop-in replacement for L/include/limu/sensors/lidar/helpers/voxel_hash_map.hpp (lidar::VoxelHashMap, :14-48):
// same constructor and member signatures, the map itself lives in B200 HBM behind liblimu_cuda's C ABI.
// Put include/limu_dropin BEFORE the reference's include directory on the compiler's search path.
//
// Differences a caller can observe:
//   * the public `tsl::robin_map ... map` member (:40) does not exist (a device table cannot be exposed as
//     a host container); use size() / voxels() / pointcloud() instead;
//   * failures (no GPU, voxel index out of the packed key range, out of memory) throw std::runtime_error
//     -- the reference has no error path at all here.
#ifndef VOXEL_HASH_MAP_HPP
#define VOXEL_HASH_MAP_HPP

#include <tuple>
#include <utility>
#include <vector>

#include "common.hpp"   // the reference's own types: utils::Vec3d, utils::Vec3dVector, utils::Vec3_Vec3Tuple
#include "limu_dropin/runtime.hpp"

namespace lidar
{
    using SE3d = Sophus::SE3d;
    class VoxelHashMap
    {
    public:
        VoxelHashMap(double vox_size, double max_distance, int max_points_per_voxel, int vox_side_length = 3)
            : cap_(max_points_per_voxel)
        {
            (void)vox_side_length;   // stored as vox_cube and never read in the reference either (hpp:19,47)
            limu_dropin::check(limu_map_create(limu_dropin::context(), vox_size, max_distance, max_points_per_voxel, 0, &h_), "VoxelHashMap");
        }
        ~VoxelHashMap() { if (h_ && owned_) limu_map_destroy(h_); }
        VoxelHashMap(const VoxelHashMap &) = delete;
        VoxelHashMap &operator=(const VoxelHashMap &) = delete;
        VoxelHashMap(VoxelHashMap &&o) noexcept : h_(o.h_), owned_(o.owned_), cap_(o.cap_) { o.h_ = nullptr; }

        // insert points into map (voxel_hash_map.cpp:12-62)
        void insert_points(const utils::Vec3dVector &points)
        {
            limu_dropin::check(limu_map_insert(h_, data(points), static_cast<int64_t>(points.size())), "insert_points");
        }

        // voxel_hash_map.cpp:64-102
        utils::Vec3d get_closest_neighbour(const utils::Vec3d &point)
        {
            utils::Vec3d out;
            limu_dropin::check(limu_map_closest(h_, point.data(), 1, out.data(), nullptr, nullptr), "get_closest_neighbour");
            return out;
        }

        // voxel_hash_map.cpp:104-130 -> {source, target}, pairs in query order
        utils::Vec3_Vec3Tuple get_correspondences(const utils::Vec3dVector &points, double max_correspondance)
        {
            utils::Vec3dVector source(points.size()), target(points.size());
            int64_t n = 0;
            limu_dropin::check(limu_map_correspondences(h_, data(points), static_cast<int64_t>(points.size()), max_correspondance,
                                                        data(source), data(target), nullptr, &n), "get_correspondences");
            source.resize(static_cast<size_t>(n));
            target.resize(static_cast<size_t>(n));
            return {source, target};
        }

        // update map points (voxel_hash_map.cpp:132-144)
        void update(const utils::Vec3dVector &points, const utils::Vec3d &origin)
        {
            limu_dropin::check(limu_map_update_origin(h_, data(points), static_cast<int64_t>(points.size()), origin.data()), "update");
        }
        void update(const utils::Vec3dVector &points, const SE3d &pose)
        {
            double p[7];
            limu_dropin::to_pose7(pose, p);
            limu_dropin::check(limu_map_update(h_, data(points), static_cast<int64_t>(points.size()), p), "update");
        }

        void remove_points_from_far(const utils::Vec3d &origin)   // voxel_hash_map.cpp:146-171 (without the self-deadlock)
        {
            limu_dropin::check(limu_map_remove_far(h_, origin.data()), "remove_points_from_far");
        }

        utils::Vec3dVector pointcloud() const   // voxel_hash_map.cpp:173-198, voxel creation order
        {
            int64_t n = 0;
            limu_dropin::check(limu_map_pointcloud(h_, nullptr, 0, &n), "pointcloud");
            utils::Vec3dVector out(static_cast<size_t>(n));
            if (n > 0) limu_dropin::check(limu_map_pointcloud(h_, out.front().data(), n, &n), "pointcloud");
            return out;
        }

        void clear() { limu_dropin::check(limu_map_clear(h_), "clear"); }
        bool empty() const
        {
            int e = 0;
            limu_dropin::check(limu_map_empty(h_, &e), "empty");
            return e != 0;
        }

        // ---- replacements for reading the public robin_map member -------------------------------------
        size_t size() const   // == map.size(): number of occupied voxels
        {
            int64_t nv = 0;
            limu_dropin::check(limu_map_size(h_, &nv, nullptr), "size");
            return static_cast<size_t>(nv);
        }
        // (voxel, points) in creation order
        std::vector<std::pair<utils::Voxel, utils::Vec3dVector>> voxels() const
        {
            int64_t nv = 0, np = 0;
            limu_dropin::check(limu_map_dump(h_, nullptr, nullptr, nullptr, 0, 0, &nv, &np), "voxels");
            std::vector<int32_t> keys(static_cast<size_t>(3 * nv + 1)), counts(static_cast<size_t>(nv + 1));
            std::vector<double> pts(static_cast<size_t>(3 * np + 1));
            limu_dropin::check(limu_map_dump(h_, keys.data(), counts.data(), pts.data(), nv, np, &nv, &np), "voxels");
            std::vector<std::pair<utils::Voxel, utils::Vec3dVector>> out;
            size_t w = 0;
            for (int64_t i = 0; i < nv; ++i) {
                utils::Vec3dVector v;
                for (int r = 0; r < counts[i]; ++r, ++w) v.emplace_back(pts[3 * w], pts[3 * w + 1], pts[3 * w + 2]);
                out.emplace_back(utils::Voxel(keys[3 * i], keys[3 * i + 1], keys[3 * i + 2]), std::move(v));
            }
            return out;
        }

        limu_map *handle() const { return h_; }
        // non-owning view over a map owned by someone else (KissICP's local map)
        VoxelHashMap(limu_map *borrowed, int cap) : h_(borrowed), owned_(false), cap_(cap) {}

    private:
        static const double *data(const utils::Vec3dVector &v) { return v.empty() ? nullptr : v.front().data(); }
        static double *data(utils::Vec3dVector &v) { return v.empty() ? nullptr : v.front().data(); }
        limu_map *h_ = nullptr;
        bool owned_ = true;
        int cap_;
    };
}
#endif
