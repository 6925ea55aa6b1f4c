// Drop-in replacement for L/include/limu/sensors/lidar/helpers/deskew.hpp:8-20 (lidar::MotionCompensator).
#ifndef DESKEW_HPP
#define DESKEW_HPP
#include <vector>

#include "common.hpp"
#include "limu_dropin/runtime.hpp"

namespace lidar
{
    using SE3d = Sophus::SE3d;
    class MotionCompensator
    {
    public:
        explicit MotionCompensator() : mid_pose_timestamp(0.5){};

        // deskew.cpp:10-28: exp((t_i - 0.5) * log(start^-1 * end)) * p_i
        utils::Vec3dVector deskew_scan(
            const utils::PointCloudXYZI &frame, const std::vector<double> &timestamps,
            const SE3d &start_pose, const SE3d &end_pose)
        {
            // the PCL records and the FP64 timestamps go to the device as they are (no host repack)
            const size_t n = frame.points.size();
            double T0[7], T1[7];
            limu_dropin::to_pose7(start_pose, T0);
            limu_dropin::to_pose7(end_pose, T1);
            utils::Vec3dVector out(n);
            limu_dropin::check(limu_deskew_cloud(limu_dropin::context(), n ? &frame.points[0] : nullptr, static_cast<int32_t>(sizeof(frame.points[0])),
                                                 timestamps.data(), static_cast<int64_t>(n), T0, T1, n ? out.front().data() : nullptr), "deskew_scan");
            return out;
        }

    private:
        double mid_pose_timestamp;
    };
}
#endif
