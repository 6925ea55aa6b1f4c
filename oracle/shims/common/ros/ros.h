// Oracle shim: just enough of roscpp for the reference headers that the path includes
// transitively (icp.hpp:25 -> lidar/frame.hpp). No ROS runtime; params return their defaults.
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
// roscpp's own headers (ros/time.h, ros/duration.h -> <math.h>) put the C math names into the global namespace; the reference's
// lidar/frame.cpp:144,161 relies on that for its unqualified isnan() / atan2() (libstdc++'s <math.h>: `using std::isnan;` etc.,
// so atan2(float, float) is the FLOAT overload).
#include <math.h>
namespace ros {
struct Time {
    double t = 0;
    Time() = default;
    explicit Time(double s) : t(s) {}
    double toSec() const { return t; }
    Time &fromSec(double s) { t = s; return *this; }
    static Time now() { return Time(0); }
};
struct Duration { double d = 0; explicit Duration(double s = 0) : d(s) {} double toSec() const { return d; } };
class NodeHandle {
public:
    template <class T> bool param(const std::string &, T &v, const T &d) const { v = d; return false; }
    template <class T> T param(const std::string &, const T &d) const { return d; }
};
}  // namespace ros
#define ROS_INFO(...) do { std::printf(__VA_ARGS__); std::printf("\n"); } while (0)
#define ROS_WARN(...) do { std::printf(__VA_ARGS__); std::printf("\n"); } while (0)
#define ROS_ERROR(...) do { std::printf(__VA_ARGS__); std::printf("\n"); } while (0)
#define ROS_INFO_STREAM(x)
#define ROS_WARN_STREAM(x)
