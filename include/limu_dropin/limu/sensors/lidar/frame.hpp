// Drop-in replacement for L/include/limu/sensors/lidar/frame.hpp (frame::Lidar, :28-168) + L/src/sensors/lidar/frame.cpp:
// same constructor (ros::NodeHandle&), ProcessingInfo fields, initialize() / process_frame() and buffer accessors; the
// per-point work of process_frame (range gate, offset time, time sort, split, timestamp normalisation -- frame.cpp:101-193,
// :28-99) runs on the B200 behind limu_preprocess_frame. Header-only: the reference's frame.cpp is not
// compiled when this header shadows its own.
//
// Differences a caller can observe:
//   * points of EQUAL offset time keep message order (the reference's std::sort leaves their order unspecified);
//   * a ring index >= num_scan_lines in the constant-rotation path throws std::runtime_error (the reference indexes its
//     per-ring vectors out of bounds); so does an empty message (the reference dereferences max_element of an empty vector).
#ifndef LIDAR_FRAME_HPP
#define LIDAR_FRAME_HPP

#include <pcl_conversions/pcl_conversions.h>
#include "geometry_msgs/TransformStamped.h"
#include "sensor_msgs/PointCloud2.h"
#include "nav_msgs/Odometry.h"
#include "common.hpp"
#include <ros/ros.h>

#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "limu_dropin/runtime.hpp"

// the message layout the reference registers with PCL (lidar/frame.hpp:12-23); kept so code that names LidarPoint still compiles
struct EIGEN_ALIGN16 LidarPoint
{
    PCL_ADD_POINT4D;
    std::uint8_t intensity;
    std::uint16_t ring;
    double timestamp;
    EIGEN_MAKE_ALIGNED_OPERATOR_NEW
};

POINT_CLOUD_REGISTER_POINT_STRUCT(
    LidarPoint,
    (float, x, x)(float, y, y)(float, z, z)(std::uint8_t, intensity, intensity)(std::uint16_t, ring, ring)(double, timestamp, timestamp))

namespace frame
{
    using PointCloud = utils::PointCloudXYZI;
    class Lidar
    {
    public:
        typedef std::shared_ptr<Lidar> Ptr;

        struct ProcessingInfo   // lidar/frame.hpp:34-58: the fields are the interface KissICP and the launch files read
        {
            typedef std::shared_ptr<ProcessingInfo> Ptr;
            double frame_rate, max_range, min_range, min_angle, max_angle;
            int num_scan_lines, frame_split_num;
            double voxel_size;
            int vox_side_length, max_points_per_voxel;
            bool deskew;
            double min_motion_th;
            int icp_max_iteration;
            double initial_threshold, estimation_threshold;
        };

        explicit Lidar(ros::NodeHandle &nh) : config(std::make_shared<ProcessingInfo>())
        {
            // ROS parameter names and defaults of lidar/frame.hpp:64-80, read through two small typed helpers
            ProcessingInfo &c = *config;
            const auto real = [&nh](const char *key, double &dst, double fallback) { nh.param<double>(key, dst, fallback); };
            const auto whole = [&nh](const char *key, int &dst, int fallback) { nh.param<int>(key, dst, fallback); };
            real("frame_rate", c.frame_rate, 10.0);
            real("max_range", c.max_range, 100.0);
            real("min_range", c.min_range, 5.0);
            real("min_angle", c.min_angle, 0.0);
            real("max_angle", c.max_angle, 360.0);
            whole("num_scan_lines", c.num_scan_lines, 16);
            whole("frame_split_num", c.frame_split_num, 1);
            real("voxel_size", c.voxel_size, c.max_range / 100.0);   // default: one hundredth of the range
            whole("vox_side_length", c.vox_side_length, 3);
            whole("max_points_per_voxel", c.max_points_per_voxel, 10);
            nh.param<bool>("deskew", c.deskew, false);
            real("min_motion_th", c.min_motion_th, 0.1);
            whole("icp_max_iteration", c.icp_max_iteration, 500);
            real("initial_threshold", c.initial_threshold, 2.0);
            real("estimation_threshold", c.estimation_threshold, 1e-4);
        }

        // frame.cpp:10-26: count the message, detect a looped bag, hold the message for process_frame
        void initialize(const sensor_msgs::PointCloud2::ConstPtr &msg)
        {
            std::lock_guard<std::mutex> lock(data_mutex);
            ++scan_count;
            const double t = msg->header.stamp.toSec();
            if (t < prev_timestamp) {
                ROS_ERROR("Lidar Buffer Looped. Clearning buffer");
                processed_buffer.clear(); timestamps.clear(); accumulated_segment_time.clear();
            }
            msg_holder = msg;
            prev_timestamp = t;
        }

        // frame.cpp:101-193 (+ sort_clouds :28-51, split_clouds :53-99) on the device; pushes the segments to the three deques
        void process_frame()
        {
            std::lock_guard<std::mutex> lock(data_mutex);
            const sensor_msgs::PointCloud2 &m = *msg_holder;
            std::string names;
            std::vector<int32_t> offs, types, counts;
            for (const auto &f : m.fields) {
                names.append(f.name); names.push_back('\0');
                offs.push_back(static_cast<int32_t>(f.offset)); types.push_back(static_cast<int32_t>(f.datatype)); counts.push_back(static_cast<int32_t>(f.count));
            }
            limu_cloud_fields cf;
            if (limu_cloud_fields_from_pointfields(static_cast<int32_t>(m.fields.size()), names.data(), offs.data(), types.data(), counts.data(),
                                                   static_cast<int32_t>(m.point_step), &cf) != LIMU_OK)
                throw std::runtime_error(limu_last_error());   // "Field 't', 'timestamp' or 'time' not existing" (calculation_helpers.cpp:13-16)
            limu_lidar_config lc;
            lc.min_range = config->min_range; lc.max_range = config->max_range; lc.min_angle = config->min_angle; lc.max_angle = config->max_angle;
            lc.frame_rate = config->frame_rate; lc.num_scan_lines = config->num_scan_lines; lc.frame_split_num = config->frame_split_num;
            const int64_t n = static_cast<int64_t>(m.height) * m.width;
            if (n <= 0) throw std::runtime_error("frame::Lidar::process_frame: empty point cloud message");
            static_assert(sizeof(utils::PointNormal) == 48, "pcl::PointXYZINormal is 48 bytes");
            std::vector<utils::PointNormal> rec(static_cast<size_t>(n));
            std::vector<double> ts(static_cast<size_t>(n));
            const int32_t max_seg = 64;
            int64_t sizes[64];
            double seg_time[64];
            int32_t nseg = 0;
            limu_dropin::check(limu_preprocess_frame(limu_dropin::context(), m.data.data(), n, &cf, &lc, m.header.stamp.toSec(), scan_count,
                                                     rec.data(), ts.data(), max_seg, sizes, seg_time, &nseg), "process_frame");
            size_t at = 0;
            for (int32_t k = 0; k < nseg; ++k) {
                const size_t sz = static_cast<size_t>(sizes[k]);
                PointCloud::Ptr cloud(new PointCloud());
                cloud->points.assign(rec.begin() + at, rec.begin() + at + sz);
                processed_buffer.emplace_back(std::move(cloud));
                timestamps.emplace_back(ts.begin() + at, ts.begin() + at + sz);
                accumulated_segment_time.push_back(seg_time[k]);
                at += sz;
            }
        }

        // accessors (lidar/frame.hpp:89-127)
        double return_prev_ts() { std::unique_lock<std::mutex> lock(data_mutex); return prev_timestamp; }
        bool buffer_empty() { std::unique_lock<std::mutex> lock(data_mutex); return processed_buffer.empty(); }
        PointCloud::Ptr get_lidar_buffer_front() { std::unique_lock<std::mutex> lock(data_mutex); return processed_buffer.front(); }
        std::vector<double> get_segment_ts_front() { std::unique_lock<std::mutex> lock(data_mutex); return timestamps.front(); }
        double curr_acc_segment_time() { std::unique_lock<std::mutex> lock(data_mutex); return accumulated_segment_time.front(); }
        void pop()
        {
            std::unique_lock<std::mutex> lock(data_mutex);
            timestamps.pop_front(); processed_buffer.pop_front(); accumulated_segment_time.pop_front();
        }

        // frame.cpp:195-221: fill the two outgoing ROS messages
        void set_current_pose_nav(const utils::Vec3d &translation, const Eigen::Quaterniond &quat, const ros::Time &time, std::string &odom_frame,
                                  std::string &child_frame)
        {
            current_pose.header.stamp = time; current_pose.header.frame_id = odom_frame; current_pose.child_frame_id = child_frame;
            auto &r = current_pose.transform.rotation;
            r.x = quat.x(); r.y = quat.y(); r.z = quat.z(); r.w = quat.w();
            auto &t = current_pose.transform.translation;
            t.x = translation.x(); t.y = translation.y(); t.z = translation.z();
            odom_msg.header.stamp = time; odom_msg.header.frame_id = odom_frame; odom_msg.child_frame_id = child_frame;
            auto &o = odom_msg.pose.pose.orientation;
            o.x = quat.x(); o.y = quat.y(); o.z = quat.z(); o.w = quat.w();
            auto &p = odom_msg.pose.pose.position;
            p.x = translation.x(); p.y = translation.y(); p.z = translation.z();
        }

    public:
        geometry_msgs::TransformStamped current_pose;
        nav_msgs::Odometry odom_msg;
        std::shared_ptr<ProcessingInfo> config;

    private:
        sensor_msgs::PointCloud2::ConstPtr msg_holder;
        std::deque<std::vector<double>> timestamps;
        std::deque<double> accumulated_segment_time;
        std::deque<PointCloud::Ptr> processed_buffer;
        std::mutex data_mutex;
        double prev_timestamp = 0.0;
        int scan_count = 0;
    };
}
#endif
