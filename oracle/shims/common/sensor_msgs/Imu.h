// Oracle shim: sensor_msgs::Imu as a plain struct.
#pragma once
#include <memory>
#include "PointCloud2.h"
namespace geometry_msgs {
struct Vector3 { double x = 0, y = 0, z = 0; };
struct Point { double x = 0, y = 0, z = 0; };
struct Quaternion { double x = 0, y = 0, z = 0, w = 1; };
}  // namespace geometry_msgs
namespace sensor_msgs {
struct Imu {
    using Ptr = std::shared_ptr<Imu>;
    using ConstPtr = std::shared_ptr<const Imu>;
    std_msgs::Header header;
    geometry_msgs::Quaternion orientation;
    geometry_msgs::Vector3 angular_velocity, linear_acceleration;
};
}  // namespace sensor_msgs
