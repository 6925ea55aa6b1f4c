// Oracle shim: pcl_conversions reduced to what frame::Lidar::process_frame needs (L/src/sensors/lidar/frame.cpp:107):
// pcl::fromROSMsg copies the message fields whose NAME matches a member of the point struct into that member, the way
// PCL's field map does for a POINT_CLOUD_REGISTER_POINT_STRUCT'ed type (pcl/conversions.h createMapping/fromPCLPointCloud2:
// name + datatype must match, unmatched members stay value-initialised). The shim knows the member names the reference
// registers (lidar/frame.hpp:21-23: x, y, z, intensity, ring, timestamp) and detects which of them a point type has.
#pragma once
#include <cstring>
#include <type_traits>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <sensor_msgs/PointCloud2.h>
namespace pcl {
namespace shim_detail {
template <class...> using void_t = void;
#define LIMU_SHIM_HAS(member)                                                                          \
    template <class T, class = void> struct has_##member : std::false_type {};                         \
    template <class T> struct has_##member<T, void_t<decltype(std::declval<T &>().member)>> : std::true_type {};
LIMU_SHIM_HAS(x) LIMU_SHIM_HAS(y) LIMU_SHIM_HAS(z) LIMU_SHIM_HAS(intensity) LIMU_SHIM_HAS(ring) LIMU_SHIM_HAS(timestamp)
#undef LIMU_SHIM_HAS
template <class M> std::uint8_t datatype_of() {
    using F = sensor_msgs::PointField;
    return std::is_same<M, float>::value ? F::FLOAT32 : std::is_same<M, double>::value ? F::FLOAT64 : std::is_same<M, std::uint8_t>::value ? F::UINT8
         : std::is_same<M, std::uint16_t>::value ? F::UINT16 : std::is_same<M, std::uint32_t>::value ? F::UINT32 : std::is_same<M, std::int8_t>::value ? F::INT8
         : std::is_same<M, std::int16_t>::value ? F::INT16 : F::INT32;
}
template <class M> void load(const sensor_msgs::PointCloud2 &msg, const char *name, std::size_t i, M &dst) {
    for (const auto &f : msg.fields)
        if (f.name == name && f.datatype == datatype_of<M>()) { std::memcpy(&dst, msg.data.data() + i * msg.point_step + f.offset, sizeof(M)); return; }
}
}  // namespace shim_detail
template <class T> void fromROSMsg(const sensor_msgs::PointCloud2 &msg, PointCloud<T> &cloud) {
    const std::size_t n = static_cast<std::size_t>(msg.width) * msg.height;
    cloud.points.assign(n, T());
    cloud.width = msg.width; cloud.height = msg.height; cloud.is_dense = msg.is_dense;
    for (std::size_t i = 0; i < n; ++i) {
        T &p = cloud.points[i];
        if constexpr (shim_detail::has_x<T>::value) shim_detail::load(msg, "x", i, p.x);
        if constexpr (shim_detail::has_y<T>::value) shim_detail::load(msg, "y", i, p.y);
        if constexpr (shim_detail::has_z<T>::value) shim_detail::load(msg, "z", i, p.z);
        if constexpr (shim_detail::has_intensity<T>::value) shim_detail::load(msg, "intensity", i, p.intensity);
        if constexpr (shim_detail::has_ring<T>::value) shim_detail::load(msg, "ring", i, p.ring);
        if constexpr (shim_detail::has_timestamp<T>::value) shim_detail::load(msg, "timestamp", i, p.timestamp);
    }
}
template <class T> void toROSMsg(const PointCloud<T> &, sensor_msgs::PointCloud2 &) {}
}  // namespace pcl
