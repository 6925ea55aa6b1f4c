// Oracle (test infrastructure): C entry point over the reference's UNMODIFIED frame::Lidar::initialize / process_frame
// (L/src/sensors/lidar/frame.cpp:10-26, :101-193, with sort_clouds :28-51 and split_clouds :53-99) -- the host
// preprocessing that feeds KissICP::register_frame and that SURVEY section 8(f) lists as N3. The driver wraps the caller's
// bytes in the sensor_msgs::PointCloud2 shim, runs the reference, and returns the processed segments.
// `#define private public` only lifts access control (config -> setup(), scan_count, the three output deques); no
// reference source is modified or copied. Everything frame.hpp pulls in is included first so the lift touches only its text.
#include <cmath>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>
#include <pcl_conversions/pcl_conversions.h>
#include "geometry_msgs/TransformStamped.h"
#include "sensor_msgs/PointCloud2.h"
#include "nav_msgs/Odometry.h"
#include "common.hpp"
#include <ros/ros.h>
#include "limu/utils/calculation_helpers.hpp"
#define private public
#include "limu/sensors/lidar/frame.hpp"
#undef private

extern "C" {

// fields: nf entries; names = nf NUL-terminated strings back to back; offsets / datatypes (sensor_msgs::PointField codes) / counts.
// cfg = {min_range, max_range, min_angle, max_angle, frame_rate, num_scan_lines, frame_split_num}; scan_count = the value the
// reference's counter holds DURING process_frame (initialize() increments it first, frame.cpp:13).
// Outputs (capacity n points / max_segments segments): seg_sizes[k], seg_time[k] (accumulated_segment_time), and for the
// concatenated segments rec5[5*j] = {x, y, z, intensity, curvature} and ts[j] (the normalised per-point timestamps).
// Returns the number of segments pushed to processed_buffer.
long ref_process_frame(const unsigned char *data, long n, int point_step, int nf, const char *names, const int *offsets, const int *datatypes,
                       const int *counts, const double *cfg, double message_time, int scan_count, long max_segments, long *seg_sizes, double *seg_time,
                       float *rec5, double *ts) {
    if (n <= 0 || (int)cfg[6] < 1) return 0;   // the reference dereferences max_element of an empty vector / divides by zero here
    ros::NodeHandle nh;
    frame::Lidar lidar(nh);
    lidar.config->min_range = cfg[0]; lidar.config->max_range = cfg[1]; lidar.config->min_angle = cfg[2]; lidar.config->max_angle = cfg[3];
    lidar.config->frame_rate = cfg[4]; lidar.config->num_scan_lines = (int)cfg[5]; lidar.config->frame_split_num = (int)cfg[6];
    lidar.scan_ang_vel = utils::calc_scan_ang_vel(lidar.config->frame_rate);   // lidar/frame.hpp:82-83, after the parameters are in place
    lidar.setup();
    lidar.scan_count = scan_count - 1;
    auto msg = std::make_shared<sensor_msgs::PointCloud2>();
    msg->header.stamp.fromSec(message_time);
    msg->height = 1; msg->width = (std::uint32_t)n; msg->point_step = (std::uint32_t)point_step; msg->row_step = (std::uint32_t)(point_step * n);
    msg->data.assign(data, data + (size_t)n * point_step);
    const char *p = names;
    for (int k = 0; k < nf; ++k) {
        sensor_msgs::PointField f;
        f.name = p; p += f.name.size() + 1;
        f.offset = (std::uint32_t)offsets[k]; f.datatype = (std::uint8_t)datatypes[k]; f.count = (std::uint32_t)counts[k];
        msg->fields.push_back(f);
    }
    lidar.initialize(msg);      // frame.cpp:10-26 (scan_count++, msg_holder)
    try {
        lidar.process_frame();  // frame.cpp:101-193
    } catch (const std::runtime_error &) {
        return -1;              // "Field 't', 'timestamp' or 'time' not existing" (calculation_helpers.cpp:13-16)
    }
    long nseg = 0, at = 0;
    while (!lidar.processed_buffer.empty() && nseg < max_segments) {
        const auto cloud = lidar.processed_buffer.front();
        const auto &t = lidar.timestamps.front();
        seg_sizes[nseg] = (long)cloud->points.size();
        seg_time[nseg] = lidar.accumulated_segment_time.front();
        for (size_t j = 0; j < cloud->points.size(); ++j, ++at) {
            const auto &q = cloud->points[j];
            rec5[5 * at] = q.x; rec5[5 * at + 1] = q.y; rec5[5 * at + 2] = q.z; rec5[5 * at + 3] = q.intensity; rec5[5 * at + 4] = q.curvature;
            ts[at] = t[j];
        }
        lidar.processed_buffer.pop_front(); lidar.timestamps.pop_front(); lidar.accumulated_segment_time.pop_front();
        ++nseg;
    }
    return nseg;
}

}  // extern "C"
