// Drop-in replacement for L/include/limu/sensors/lidar/icp.hpp:31-68 (lidar::KissICP): same constructor
// (from frame::Lidar::ProcessingInfo::Ptr) and member functions; the pipeline state (local map, pose
// history, adaptive threshold) lives behind one limu_odom handle.
#ifndef KISS_ICP_HPP
#define KISS_ICP_HPP

#include "sensor_msgs/PointCloud2.h"
#include "helpers/registration.hpp"
#include "helpers/voxel_hash_map.hpp"
#include "helpers/deskew.hpp"
#include "common.hpp"
#include "limu/sensors/lidar/frame.hpp"   // the reference's own frame::Lidar::ProcessingInfo (lidar/frame.hpp:34-58)

namespace lidar
{
    using SE3d = Sophus::SE3d;
    using ReturnTuple = std::tuple<utils::Vec3dVector, utils::Vec3dVector, SE3d>;
    class KissICP
    {
    public:
        typedef std::shared_ptr<KissICP> Ptr;
        explicit KissICP(const frame::Lidar::ProcessingInfo::Ptr &config)
            : config(*config), scan_duration(1 / (double(config->frame_split_num) * config->frame_rate))
        {
            limu_odom_config c;
            limu_odom_default_config(&c);
            c.voxel_size = config->voxel_size; c.max_range = config->max_range;
            c.max_points_per_voxel = config->max_points_per_voxel; c.deskew = config->deskew ? 1 : 0;
            c.min_motion_th = config->min_motion_th; c.icp_max_iteration = config->icp_max_iteration;
            c.initial_threshold = config->initial_threshold; c.estimation_threshold = config->estimation_threshold;
            limu_dropin::check(limu_odom_create(limu_dropin::context(), &c, &h_), "KissICP");
        }
        ~KissICP() { if (h_) limu_odom_destroy(h_); }
        KissICP(const KissICP &) = delete;
        KissICP &operator=(const KissICP &) = delete;

        // outlier removal from points (icp.cpp:88-124)
        utils::Vec3dVector iqr_processing(const utils::Vec3dVector &frame)
        {
            utils::Vec3dVector out(frame.size());
            int64_t n = 0;
            limu_dropin::check(limu_iqr_filter(limu_dropin::context(), ptr(frame), static_cast<int64_t>(frame.size()), ptr(out), &n, nullptr), "iqr_processing");
            out.resize(static_cast<size_t>(n));
            return out;
        }

        // icp.cpp:36-47 (deskew gate: config.deskew && poses.size() > 2)
        utils::Vec3dVector deskew_scan(const utils::PointCloudXYZI &frame, const std::vector<double> &timestamps)
        {
            const auto poses = poses_();
            if (!config.deskew || poses.size() <= 2) {
                utils::Vec3dVector e(frame.points.size());
                for (size_t i = 0; i < e.size(); ++i) e[i] = utils::Vec3d(frame.points[i].x, frame.points[i].y, frame.points[i].z);
                return e;
            }
            return compensator.deskew_scan(frame, timestamps, poses[poses.size() - 2], poses[poses.size() - 1]);
        }

        // register frame (icp.cpp:49-55): raw cloud + normalised timestamps in, one H2D copy, one synchronisation
        ReturnTuple register_frame(const utils::PointCloudXYZI &pointcloud, const std::vector<double> &timestamps)
        {
            // the PCL records and the FP64 timestamps go to the device as they are (no host repack)
            const size_t n = pointcloud.points.size();
            utils::Vec3dVector down(n), src(n);
            int64_t nd = 0, ns = 0;
            double pose[7];
            limu_dropin::check(limu_odom_register_cloud(h_, n ? &pointcloud.points[0] : nullptr, static_cast<int32_t>(sizeof(pointcloud.points[0])), timestamps.data(),
                                                        static_cast<int64_t>(n), pose, ptr(down), &nd, ptr(src), &ns, &stats_), "register_frame");
            down.resize(static_cast<size_t>(nd));
            src.resize(static_cast<size_t>(ns));
            return {down, src, limu_dropin::from_pose7(pose)};
        }
        // Extension (no counterpart in the reference): a caller that already holds the NEXT cloud -- odom_run.cpp pops frames from a buffer --
        // may announce it, so that its upload (and, by default, its deskew + downsampling) overlaps the registration of the current one.
        // Both containers must stay untouched until they are passed to register_frame.
        void prefetch(const utils::PointCloudXYZI &pointcloud, const std::vector<double> &timestamps)
        {
            const size_t n = pointcloud.points.size();
            if (n) limu_dropin::check(limu_odom_prefetch_cloud(h_, &pointcloud.points[0], static_cast<int32_t>(sizeof(pointcloud.points[0])), timestamps.data(), static_cast<int64_t>(n)), "prefetch");
        }
        // icp.cpp:58-86
        ReturnTuple register_frame(const utils::Vec3dVector &frame)
        {
            const size_t n = frame.size();
            utils::Vec3dVector down(n), src(n);
            int64_t nd = 0, ns = 0;
            double pose[7];
            limu_dropin::check(limu_odom_register_points(h_, ptr(frame), static_cast<int64_t>(n), pose, ptr(down), &nd, ptr(src), &ns, &stats_), "register_frame");
            down.resize(static_cast<size_t>(nd));
            src.resize(static_cast<size_t>(ns));
            return {down, src, limu_dropin::from_pose7(pose)};
        }

        // downsample pointcloud KISS-ICP Downsampling scheme (icp.cpp:126-136) -> {source, frame_downsample}
        utils::Vec3_Vec3Tuple voxelize(const utils::Vec3dVector &frame, const double vox_size)
        {
            utils::Vec3dVector src(frame.size()), down(frame.size());
            int64_t ns = 0, nd = 0;
            limu_dropin::check(limu_voxelize(limu_dropin::context(), ptr(frame), static_cast<int64_t>(frame.size()), vox_size, ptr(src), &ns, ptr(down), &nd), "voxelize");
            src.resize(static_cast<size_t>(ns));
            down.resize(static_cast<size_t>(nd));
            return {src, down};
        }

        SE3d get_prediction_model() const   // icp.cpp:146-154
        {
            double p[7];
            limu_dropin::check(limu_odom_prediction(h_, p), "get_prediction_model");
            return limu_dropin::from_pose7(p);
        }
        double get_adaptive_threshold()     // icp.cpp:138-144 (accumulates, like the reference)
        {
            double s = 0;
            limu_dropin::check(limu_odom_adaptive_threshold(h_, &s), "get_adaptive_threshold");
            return s;
        }
        utils::Vec3Tuple current_vel()      // icp.cpp:165-172
        {
            if (!has_moved()) return {utils::Vec3d::Zero(), utils::Vec3d::Zero()};
            const auto poses = poses_();
            const std::size_t N = poses.size();
            const utils::vector<6> twist = (poses[N - 2].inverse() * poses[N - 1]).log() / scan_duration;
            return {twist.head<3>(), twist.tail<3>()};
        }
        bool has_moved()                    // icp.cpp:156-163
        {
            int m = 0;
            limu_dropin::check(limu_odom_has_moved(h_, &m), "has_moved");
            return m != 0;
        }

        utils::Vec3dVector local_map_() const { return VoxelHashMap(limu_odom_map(h_), config.max_points_per_voxel).pointcloud(); }
        std::vector<SE3d> poses_() const
        {
            int64_t n = 0;
            limu_dropin::check(limu_odom_num_poses(h_, &n), "poses_");
            std::vector<SE3d> out(static_cast<size_t>(n));
            for (int64_t i = 0; i < n; ++i) {
                double p[7];
                limu_dropin::check(limu_odom_pose(h_, i, p), "poses_");
                out[static_cast<size_t>(i)] = limu_dropin::from_pose7(p);
            }
            return out;
        }
        const limu_frame_stats &last_stats() const { return stats_; }   // extension: iterations, sigma, counts of the last frame

    private:
        static const double *ptr(const utils::Vec3dVector &v) { return v.empty() ? nullptr : v.front().data(); }
        static double *ptr(utils::Vec3dVector &v) { return v.empty() ? nullptr : v.front().data(); }
        frame::Lidar::ProcessingInfo config;
        MotionCompensator compensator;
        double scan_duration;
        limu_odom *h_ = nullptr;
        limu_frame_stats stats_{};
    };
}
#endif
