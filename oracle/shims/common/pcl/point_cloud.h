// Oracle shim: pcl::PointCloud<T> reduced to the members the reference touches.
#pragma once
#include <cstddef>
#include <cstdint>
#include <memory>
#include <string>
#include <vector>
namespace pcl {
struct PCLHeader { std::uint32_t seq = 0; std::uint64_t stamp = 0; std::string frame_id; };
template <class T>
class PointCloud {
public:
    using Ptr = std::shared_ptr<PointCloud<T>>;
    using ConstPtr = std::shared_ptr<const PointCloud<T>>;
    using PointType = T;
    PCLHeader header;
    std::vector<T> points;
    std::uint32_t width = 0, height = 1;
    bool is_dense = true;
    std::size_t size() const { return points.size(); }
    bool empty() const { return points.empty(); }
    void reserve(std::size_t n) { points.reserve(n); }
    void resize(std::size_t n) { points.resize(n); }
    void clear() { points.clear(); }
    void push_back(const T &p) { points.push_back(p); }
    T &operator[](std::size_t i) { return points[i]; }
    const T &operator[](std::size_t i) const { return points[i]; }
    typename std::vector<T>::iterator begin() { return points.begin(); }
    typename std::vector<T>::iterator end() { return points.end(); }
    typename std::vector<T>::const_iterator begin() const { return points.begin(); }
    typename std::vector<T>::const_iterator end() const { return points.end(); }
    T &back() { return points.back(); }
    T &front() { return points.front(); }
    PointCloud &operator+=(const PointCloud &o) { points.insert(points.end(), o.points.begin(), o.points.end()); return *this; }
};
}  // namespace pcl
