// Oracle shim: nav_msgs::Odometry as a plain struct.
#pragma once
#include <geometry_msgs/TransformStamped.h>
namespace nav_msgs {
struct Odometry { std_msgs::Header header; std::string child_frame_id; geometry_msgs::PoseWithCovariance pose; geometry_msgs::TwistWithCovariance twist; };
struct Path { std_msgs::Header header; std::vector<geometry_msgs::PoseStamped> poses; };
}  // namespace nav_msgs
