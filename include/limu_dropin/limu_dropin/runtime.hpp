// limu_dropin/runtime.hpp -- glue shared by the drop-in headers: one lazily created GPU context per
// process (the reference's odometry runs on a single thread, L/src/odom_run.cpp:154-185), status -> exception,
// Sophus::SE3d <-> double[7].
#ifndef LIMU_DROPIN_RUNTIME_HPP
#define LIMU_DROPIN_RUNTIME_HPP

#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>

#include <sophus/se3.hpp>

#include "limu_cuda.h"

namespace limu_dropin {

inline void check(int status, const char *what) {
    if (status != LIMU_OK) throw std::runtime_error(std::string(what) + ": " + limu_last_error());
}

// Process-wide context on device $LIMU_DEVICE (default 0). There is no CPU fallback: this throws without a B200.
inline limu_ctx *context() {
    static limu_ctx *ctx = [] {
        const char *e = std::getenv("LIMU_DEVICE");
        limu_ctx *c = nullptr;
        check(limu_ctx_create(e ? std::atoi(e) : 0, &c), "limu_ctx_create");
        return c;
    }();
    return ctx;
}

// Sophus::SE3d stores {quaternion x,y,z,w, translation}: the same 7 doubles as the C ABI's pose.
inline void to_pose7(const Sophus::SE3d &T, double p[7]) { std::memcpy(p, T.data(), 7 * sizeof(double)); }
inline Sophus::SE3d from_pose7(const double p[7]) {
    Sophus::SE3d T;
    std::memcpy(T.data(), p, 7 * sizeof(double));
    return T;
}

}  // namespace limu_dropin
#endif
