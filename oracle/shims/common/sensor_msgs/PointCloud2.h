// Oracle shim: sensor_msgs::PointCloud2 / PointField as plain structs (include satisfaction only).
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>
#include <ros/ros.h>
namespace std_msgs { struct Header { std::uint32_t seq = 0; ros::Time stamp; std::string frame_id; }; }
namespace sensor_msgs {
struct PointField {
    enum { INT8 = 1, UINT8, INT16, UINT16, INT32, UINT32, FLOAT32, FLOAT64 };
    std::string name; std::uint32_t offset = 0; std::uint8_t datatype = 0; std::uint32_t count = 0;
};
struct PointCloud2 {
    using Ptr = std::shared_ptr<PointCloud2>;
    using ConstPtr = std::shared_ptr<const PointCloud2>;
    std_msgs::Header header;
    std::uint32_t height = 0, width = 0;
    std::vector<PointField> fields;
    bool is_bigendian = false;
    std::uint32_t point_step = 0, row_step = 0;
    std::vector<std::uint8_t> data;
    bool is_dense = true;
};
}  // namespace sensor_msgs
