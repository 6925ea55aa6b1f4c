// Oracle shim: the PCL point structs the reference names (types.hpp:36-37). Field names and the
// 48-byte, 16-aligned layout of PointXYZINormal follow PCL; only the members the path reads exist.
#pragma once
#include <cstdint>
#define PCL_ADD_POINT4D float x, y, z, data_pad_;
#define POINT_CLOUD_REGISTER_POINT_STRUCT(...)
namespace pcl {
struct alignas(16) PointXYZINormal {
    float x = 0, y = 0, z = 0, data_pad_ = 1.f;
    float normal_x = 0, normal_y = 0, normal_z = 0, normal_pad_ = 0;
    float intensity = 0, curvature = 0, pad2_[2] = {0, 0};
};
static_assert(sizeof(PointXYZINormal) == 48, "PointXYZINormal layout");
struct alignas(16) PointXYZRGB {
    float x = 0, y = 0, z = 0, data_pad_ = 1.f;
    std::uint8_t b = 0, g = 0, r = 0, a = 255;
    float pad_[3] = {0, 0, 0};
};
}  // namespace pcl
