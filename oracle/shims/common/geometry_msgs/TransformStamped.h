// Oracle shim: geometry_msgs::TransformStamped as a plain struct.
#pragma once
#include <sensor_msgs/Imu.h>
namespace geometry_msgs {
struct Transform { Vector3 translation; Quaternion rotation; };
struct TransformStamped { std_msgs::Header header; std::string child_frame_id; Transform transform; };
struct Pose { Point position; Quaternion orientation; };
struct PoseWithCovariance { Pose pose; double covariance[36] = {0}; };
struct Twist { Vector3 linear, angular; };
struct TwistWithCovariance { Twist twist; double covariance[36] = {0}; };
struct PoseStamped { std_msgs::Header header; Pose pose; };
}  // namespace geometry_msgs
