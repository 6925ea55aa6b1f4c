// Oracle shim: strided field iterator over PointCloud2::data.
#pragma once
#include <cstring>
#include <stdexcept>
#include "PointCloud2.h"
namespace sensor_msgs {
template <class T>
class PointCloud2ConstIterator {
public:
    PointCloud2ConstIterator(const PointCloud2 &m, const std::string &field) : step_(m.point_step) {
        for (const auto &f : m.fields)
            if (f.name == field) { p_ = m.data.data() + f.offset; return; }
        throw std::runtime_error("Field " + field + " does not exist");
    }
    T operator*() const { T v; std::memcpy(&v, p_, sizeof(T)); return v; }
    PointCloud2ConstIterator &operator++() { p_ += step_; return *this; }
private:
    const std::uint8_t *p_ = nullptr;
    std::uint32_t step_;
};
}  // namespace sensor_msgs
