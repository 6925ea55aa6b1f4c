// Oracle (test infrastructure, not product code): C entry points over the UNMODIFIED reference
// classes, compiled from the sources where they lie under /root/reference by oracle/Makefile into
// oracle/_ref/liblimu_ref.so (serial shims: the parity oracle) and liblimu_ref_mt.so (thread-pool
// TBB shim: the timed CPU baseline). Nothing in the product links or loads this file.
//
// Pose convention on this boundary: double[7] = {qx, qy, qz, qw, tx, ty, tz}, i.e. Sophus::SE3d's
// own parameter order (sophus/se3.hpp data()).
//
// Each function below is a direct call into the reference function named beside it.
#include <cstring>
#include <memory>
#include <vector>

#include "limu/sensors/lidar/icp.hpp"

using utils::Vec3d;
using utils::Vec3dVector;
using SE3 = Sophus::SE3d;

namespace {
Vec3dVector to_vec(const double *xyz, long n) {
    Vec3dVector v(static_cast<size_t>(n));
    if (n > 0) std::memcpy(v.data(), xyz, sizeof(double) * 3 * static_cast<size_t>(n));
    return v;
}
long from_vec(const Vec3dVector &v, double *out) {
    if (out && !v.empty()) std::memcpy(out, v.data(), sizeof(double) * 3 * v.size());
    return static_cast<long>(v.size());
}
SE3 to_se3(const double *p) {
    // Build without Sophus's normalising constructor when the input is already unit length is not
    // possible through the public API; SE3(Quaternion, t) normalises (so3.hpp:528-533). Inputs
    // produced by Sophus round-trip to within 1 ulp, which the tests account for.
    Eigen::Quaterniond q(p[3], p[0], p[1], p[2]);
    return SE3(q, Vec3d(p[4], p[5], p[6]));
}
SE3 to_se3_raw(const double *p) {
    // Bit-preserving load: copies the 7 parameters straight into the object (no renormalisation).
    SE3 T;
    std::memcpy(T.data(), p, sizeof(double) * 7);
    return T;
}
void from_se3(const SE3 &T, double *p) { std::memcpy(p, T.data(), sizeof(double) * 7); }

struct Kiss {
    frame::Lidar::ProcessingInfo::Ptr cfg;
    std::unique_ptr<lidar::KissICP> icp;
};
}  // namespace

extern "C" {

// ---- utils (calculation_helpers.cpp) --------------------------------------------------------
void ref_vox_index(const double *xyz, long n, double v, int *keys) {  // get_vox_index :142-147
    for (long i = 0; i < n; ++i) {
        const auto k = utils::get_vox_index(Vec3d(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]), v);
        keys[3 * i] = k[0]; keys[3 * i + 1] = k[1]; keys[3 * i + 2] = k[2];
    }
}
unsigned long ref_voxel_hash(int x, int y, int z) { return utils::VoxelHash()(utils::Voxel(x, y, z)); }  // types.hpp:43-50
void ref_transform(const double *pose7, double *xyz, long n) {  // transform_points :121-133
    auto v = to_vec(xyz, n);
    utils::transform_points(to_se3_raw(pose7), v);
    from_vec(v, xyz);
}
void ref_se3_exp(const double *x6, double *pose7) {  // vector6d_to_mat4d :116-119
    utils::vector<6> v;
    std::memcpy(v.data(), x6, sizeof(double) * 6);
    from_se3(utils::vector6d_to_mat4d(v), pose7);
}
void ref_se3_log(const double *pose7, double *x6) {
    const utils::vector<6> v = to_se3_raw(pose7).log();
    std::memcpy(x6, v.data(), sizeof(double) * 6);
}
void ref_se3_mul(const double *a7, const double *b7, double *out7) { from_se3(to_se3_raw(a7) * to_se3_raw(b7), out7); }
void ref_se3_inv(const double *a7, double *out7) { from_se3(to_se3_raw(a7).inverse(), out7); }
void ref_se3_normalized(const double *a7, double *out7) { from_se3(to_se3(a7), out7); }
void ref_delta_pose(const double *a7, const double *b7, double *x6) {  // delta_pose :99-102
    const utils::vector<6> v = utils::delta_pose(to_se3_raw(a7), to_se3_raw(b7));
    std::memcpy(x6, v.data(), sizeof(double) * 6);
}
void ref_se3_matrix(const double *a7, double *m16_rowmajor) {
    const Eigen::Matrix4d M = to_se3_raw(a7).matrix();
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) m16_rowmajor[4 * r + c] = M(r, c);
}

// ---- VoxelHashMap (voxel_hash_map.cpp) ---------------------------------------------------------
void *ref_map_create(double vox, double max_dist, int cap) { return new lidar::VoxelHashMap(vox, max_dist, cap); }
void ref_map_destroy(void *h) { delete static_cast<lidar::VoxelHashMap *>(h); }
void ref_map_clear(void *h) { static_cast<lidar::VoxelHashMap *>(h)->clear(); }
int ref_map_empty(void *h) { return static_cast<lidar::VoxelHashMap *>(h)->empty() ? 1 : 0; }
long ref_map_num_voxels(void *h) { return static_cast<long>(static_cast<lidar::VoxelHashMap *>(h)->map.size()); }
void ref_map_insert(void *h, const double *xyz, long n) {  // insert_points :12-62
    static_cast<lidar::VoxelHashMap *>(h)->insert_points(to_vec(xyz, n));
}
void ref_map_update(void *h, const double *xyz, long n, const double *pose7) {  // update :138-144
    static_cast<lidar::VoxelHashMap *>(h)->update(to_vec(xyz, n), to_se3_raw(pose7));
}
void ref_map_remove_far(void *h, const double *o) {  // remove_points_from_far :146-171
    static_cast<lidar::VoxelHashMap *>(h)->remove_points_from_far(Vec3d(o[0], o[1], o[2]));
}
void ref_map_closest(void *h, const double *xyz, long n, double *out) {  // get_closest_neighbour :64-102
    auto *m = static_cast<lidar::VoxelHashMap *>(h);
    for (long i = 0; i < n; ++i) {
        const Vec3d r = m->get_closest_neighbour(Vec3d(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]));
        out[3 * i] = r[0]; out[3 * i + 1] = r[1]; out[3 * i + 2] = r[2];
    }
}
long ref_map_correspondences(void *h, const double *xyz, long n, double tau, double *src, double *tgt) {  // :104-130
    const auto r = static_cast<lidar::VoxelHashMap *>(h)->get_correspondences(to_vec(xyz, n), tau);
    from_vec(std::get<0>(r), src);
    return from_vec(std::get<1>(r), tgt);
}
// Dump in iteration (== creation) order. keys: 3 ints per voxel; counts: points per voxel;
// pts: counts[i] points per voxel, packed back to back. Returns number of voxels; *n_pts total.
long ref_map_dump(void *h, int *keys, int *counts, double *pts, long max_vox, long max_pts, long *n_pts) {
    auto *m = static_cast<lidar::VoxelHashMap *>(h);
    long nv = 0, np = 0;
    for (const auto &kv : m->map) {
        const auto block = kv.second.get_points();
        if (nv < max_vox) {
            if (keys) { keys[3 * nv] = kv.first[0]; keys[3 * nv + 1] = kv.first[1]; keys[3 * nv + 2] = kv.first[2]; }
            if (counts) counts[nv] = static_cast<int>(block->size());
        }
        for (const auto &p : *block) {
            if (pts && np < max_pts) { pts[3 * np] = (*p)[0]; pts[3 * np + 1] = (*p)[1]; pts[3 * np + 2] = (*p)[2]; }
            ++np;
        }
        ++nv;
    }
    if (n_pts) *n_pts = np;
    return nv;
}
long ref_map_pointcloud(void *h, double *out, long max_n) {  // pointcloud :173-198
    const auto v = static_cast<lidar::VoxelHashMap *>(h)->pointcloud();
    const long n = static_cast<long>(v.size());
    if (out) std::memcpy(out, v.data(), sizeof(double) * 3 * static_cast<size_t>(n < max_n ? n : max_n));
    return n;
}

// ---- registration.cpp ----------------------------------------------------------------------------
void ref_align(const double *src, const double *tgt, long n, double th, double *pose7) {  // align_clouds :43-92
    from_se3(lidar::align_clouds(to_vec(src, n), to_vec(tgt, n), th), pose7);
}
void ref_icp(void *map, const double *xyz, long n, const double *init7, double tau, double th,
             int max_iter, double eps, double *pose7) {  // ICP :94-130
    from_se3(lidar::ICP(*static_cast<lidar::VoxelHashMap *>(map), to_vec(xyz, n), to_se3_raw(init7), tau, th,
                        max_iter, eps), pose7);
}
// The same loop unrolled through the reference's own pieces so per-iteration state is visible
// (registration.cpp:108-126 restated with reference calls only). est_trace: 7 doubles/iteration,
// ncorr_trace: correspondences/iteration. Returns iterations executed. ref_icp() is the check
// that this unrolling is bit-identical to lidar::ICP.
int ref_icp_trace(void *map, const double *xyz, long n, const double *init7, double tau, double th,
                  int max_iter, double eps, double *pose7, double *est_trace, long *ncorr_trace,
                  double *src_after /* n*3 or null: source points after the last iteration */) {
    auto &m = *static_cast<lidar::VoxelHashMap *>(map);
    const SE3 init = to_se3_raw(init7);
    if (m.empty()) { from_se3(init, pose7); return 0; }
    Vec3dVector source = to_vec(xyz, n);
    utils::transform_points(init, source);
    SE3 T_icp = SE3();
    int j = 0;
    for (; j < max_iter; ++j) {
        const auto result = m.get_correspondences(source, tau);
        auto estimate = lidar::align_clouds(std::get<0>(result), std::get<1>(result), th);
        utils::transform_points(estimate, source);
        T_icp = estimate * T_icp;
        if (est_trace) from_se3(estimate, est_trace + 7 * j);
        if (ncorr_trace) ncorr_trace[j] = static_cast<long>(std::get<0>(result).size());
        if (estimate.log().norm() < eps) { ++j; break; }
    }
    from_se3(T_icp * init, pose7);
    if (src_after) from_vec(source, src_after);
    return j;
}

// ---- deskew.cpp ------------------------------------------------------------------------------------
void ref_deskew(const float *xyz, const double *ts, long n, const double *T0, const double *T1, double *out) {  // :10-28
    utils::PointCloudXYZI cloud;
    cloud.points.resize(static_cast<size_t>(n));
    for (long i = 0; i < n; ++i) { cloud.points[i].x = xyz[3 * i]; cloud.points[i].y = xyz[3 * i + 1]; cloud.points[i].z = xyz[3 * i + 2]; }
    std::vector<double> t(ts, ts + n);
    lidar::MotionCompensator mc;
    from_vec(mc.deskew_scan(cloud, t, to_se3_raw(T0), to_se3_raw(T1)), out);
}

// ---- threshold.cpp --------------------------------------------------------------------------------
// One AdaptiveThreshold step: update_model_deviation(dev) then compute_threshold() (threshold.cpp:16-28).
void *ref_threshold_create(double init_th, double min_motion, double max_range) { return new lidar::AdaptiveThreshold(init_th, min_motion, max_range); }
void ref_threshold_destroy(void *h) { delete static_cast<lidar::AdaptiveThreshold *>(h); }
double ref_threshold_step(void *h, const double *dev7) {
    auto *a = static_cast<lidar::AdaptiveThreshold *>(h);
    a->update_model_deviation(to_se3_raw(dev7));
    return a->compute_threshold();
}

// ---- KissICP (icp.cpp) ----------------------------------------------------------------------------
void *ref_kiss_create(double voxel_size, double max_range, int cap, int deskew, double min_motion_th,
                      int icp_max_iteration, double initial_threshold, double estimation_threshold) {
    auto *k = new Kiss;
    k->cfg = std::make_shared<frame::Lidar::ProcessingInfo>();
    auto &c = *k->cfg;
    c.frame_rate = 10.0; c.max_range = max_range; c.min_range = 5.0; c.min_angle = 0.0; c.max_angle = 360.0;
    c.num_scan_lines = 16; c.frame_split_num = 1;  // lidar/frame.hpp:64-70 defaults
    c.voxel_size = voxel_size; c.vox_side_length = 3; c.max_points_per_voxel = cap;
    c.deskew = deskew != 0; c.min_motion_th = min_motion_th; c.icp_max_iteration = icp_max_iteration;
    c.initial_threshold = initial_threshold; c.estimation_threshold = estimation_threshold;
    k->icp.reset(new lidar::KissICP(k->cfg));
    return k;
}
void ref_kiss_destroy(void *h) { delete static_cast<Kiss *>(h); }
// voxel_downsample (icp.cpp:9-30) is file-local; voxelize(frame, 2s) returns ds(frame, 0.5*(2s)) as
// its second element and 0.5*(2s) == s exactly in binary floating point.
long ref_voxel_downsample(void *h, const double *xyz, long n, double s, double *out) {
    const auto r = static_cast<Kiss *>(h)->icp->voxelize(to_vec(xyz, n), 2.0 * s);
    return from_vec(std::get<1>(r), out);
}
void ref_voxelize(void *h, const double *xyz, long n, double v, double *src, long *n_src, double *down, long *n_down) {  // :126-136
    const auto r = static_cast<Kiss *>(h)->icp->voxelize(to_vec(xyz, n), v);
    *n_src = from_vec(std::get<0>(r), src);
    *n_down = from_vec(std::get<1>(r), down);
}
long ref_iqr(void *h, const double *xyz, long n, double *out) {  // iqr_processing :88-124
    return from_vec(static_cast<Kiss *>(h)->icp->iqr_processing(to_vec(xyz, n)), out);
}
// register_frame(Vec3dVector) :58-86. Buffers must hold n points each.
void ref_kiss_register_points(void *h, const double *xyz, long n, double *down, long *n_down, double *src,
                              long *n_src, double *pose7) {
    const auto r = static_cast<Kiss *>(h)->icp->register_frame(to_vec(xyz, n));
    *n_down = from_vec(std::get<0>(r), down);
    *n_src = from_vec(std::get<1>(r), src);
    from_se3(std::get<2>(r), pose7);
}
// register_frame(cloud, timestamps) :49-55 (deskew gate :36-47 inside).
void ref_kiss_register_cloud(void *h, const float *xyz, const double *ts, long n, double *down, long *n_down,
                             double *src, long *n_src, double *pose7) {
    utils::PointCloudXYZI cloud;
    cloud.points.resize(static_cast<size_t>(n));
    for (long i = 0; i < n; ++i) { cloud.points[i].x = xyz[3 * i]; cloud.points[i].y = xyz[3 * i + 1]; cloud.points[i].z = xyz[3 * i + 2]; }
    std::vector<double> t(ts, ts + n);
    const auto r = static_cast<Kiss *>(h)->icp->register_frame(cloud, t);
    if (n_down) *n_down = from_vec(std::get<0>(r), down);
    if (n_src) *n_src = from_vec(std::get<1>(r), src);
    from_se3(std::get<2>(r), pose7);
}
long ref_kiss_num_poses(void *h) { return static_cast<long>(static_cast<Kiss *>(h)->icp->poses_().size()); }
void ref_kiss_pose(void *h, long i, double *pose7) { from_se3(static_cast<Kiss *>(h)->icp->poses_()[static_cast<size_t>(i)], pose7); }
long ref_kiss_local_map(void *h, double *out, long max_n) {
    const auto v = static_cast<Kiss *>(h)->icp->local_map_();
    const long n = static_cast<long>(v.size());
    if (out) std::memcpy(out, v.data(), sizeof(double) * 3 * static_cast<size_t>(n < max_n ? n : max_n));
    return n;
}
int ref_num_threads();  // defined by the TBB shim flavour (1 for serial)

}  // extern "C"

#ifdef LIMU_REF_MT
#include <tbb/pool.h>
extern "C" int ref_num_threads() { return tbb::shim::Pool::get().size(); }
#else
extern "C" int ref_num_threads() { return 1; }
#endif
