// Oracle shim: declaration-only pcl_conversions (nothing on the path calls it).
#pragma once
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <sensor_msgs/PointCloud2.h>
namespace pcl {
template <class T> void fromROSMsg(const sensor_msgs::PointCloud2 &, PointCloud<T> &) {}
template <class T> void toROSMsg(const PointCloud<T> &, sensor_msgs::PointCloud2 &) {}
}  // namespace pcl
