// Drop-in replacement for L/include/limu/sensors/lidar/helpers/registration.hpp:10-15
// (lidar::align_clouds, lidar::ICP): same signatures; the correspondence search, residual/Jacobian,
// normal-equation reduction AND the Gauss-Newton loop run in one persistent kernel on the B200.
#ifndef REGISTRATION_HPP
#define REGISTRATION_HPP

#include "common.hpp"
#include "voxel_hash_map.hpp"

namespace lidar
{
    using SE3d = Sophus::SE3d;

    // registration.cpp:43-92
    inline SE3d align_clouds(const utils::Vec3dVector &source, const utils::Vec3dVector &target, double th)
    {
        double pose[7];
        const int64_t n = static_cast<int64_t>(source.size() < target.size() ? source.size() : target.size());
        limu_dropin::check(limu_align(limu_dropin::context(), n ? source.front().data() : nullptr, n ? target.front().data() : nullptr, n, th,
                                      nullptr, nullptr, nullptr, pose), "align_clouds");
        return limu_dropin::from_pose7(pose);
    }

    // registration.cpp:94-130
    inline SE3d ICP(VoxelHashMap &local_map, const utils::Vec3dVector &points,
                    const SE3d &init_guess, const double max_corresp_dist, const double kernel,
                    const int &icp_max_iteration, const double &est_threshold)
    {
        double init[7], pose[7];
        limu_dropin::to_pose7(init_guess, init);
        limu_dropin::check(limu_icp(local_map.handle(), points.empty() ? nullptr : points.front().data(), static_cast<int64_t>(points.size()), init,
                                    max_corresp_dist, kernel, icp_max_iteration, est_threshold, pose, nullptr, nullptr, nullptr, nullptr), "ICP");
        return limu_dropin::from_pose7(pose);
    }
}
#endif
