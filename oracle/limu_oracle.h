/* limu_oracle.h -- ORACLE: test infrastructure, not product code.
 *
 * A plain-C (C11, libm only), single-threaded restatement of the reference's LiDAR odometry hot
 * path (Oreoluwa-Se/Lidar-Imu-Slam, env_ws/src/limu, "L/" below). It exists to check the CUDA path;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * it. It is pinned against the reference's own compiled sources (oracle/_ref/liblimu_ref.so, built
 * by oracle/Makefile) by tests/test_oracle_pin.py and against the committed fixtures in
 * tests/golden/ that were generated from that library (tests/golden/make_golden.py).
 *
 * Conventions
 *   points  : double xyz[3*n], array-of-structs, as utils::Vec3dVector (L/include/limu/utils/types.hpp:19)
 *   pose    : double[7] = {qx,qy,qz,qw,tx,ty,tz} = Sophus::SE3d parameter order
 *   twist   : double[6] = {upsilon(3), omega(3)} = Sophus::SE3d::Tangent
 *   voxel   : int[3], truncation toward zero of p/v (L/src/utils/calculation_helpers.cpp:142-147)
 * Container-order semantics that the reference inherits from tsl::robin_map are DEFINED as
 * insertion order (iteration) and creation order (address tie-break); see oracle/shims/common/tsl/robin_map.h.
 */
#ifndef LIMU_ORACLE_H
#define LIMU_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

/* ---- utils: calculation_helpers.cpp ---------------------------------------------------------- */
void lo_vox_index(const double *xyz, long n, double v, int *keys);           /* :142-147 */
void lo_transform(const double *pose7, double *xyz, long n);                 /* :121-133 */
void lo_se3_exp(const double *x6, double *pose7);                            /* :116-119, sophus/se3.hpp:852-861 */
void lo_se3_log(const double *pose7, double *x6);                            /* sophus/se3.hpp:237-253 */
void lo_se3_mul(const double *a7, const double *b7, double *out7);           /* sophus/se3.hpp:302-307 */
void lo_se3_inv(const double *a7, double *out7);                             /* sophus/se3.hpp:222-225 */
void lo_delta_pose(const double *a7, const double *b7, double *x6);          /* :99-102 */

/* ---- VoxelHashMap: helpers/voxel_hash_map.cpp, voxel_block.cpp ------------------------------------- */
typedef struct lo_map lo_map;
lo_map *lo_map_create(double vox_size, double max_distance, int max_points_per_voxel);
void lo_map_destroy(lo_map *m);
void lo_map_clear(lo_map *m);                                                /* :200-204 */
int lo_map_empty(const lo_map *m);                                           /* :206-210 */
long lo_map_num_voxels(const lo_map *m);
/* Opt-in registration variants. NOT in the reference: this restatement DEFINES them and is the parity oracle of the
 * CUDA path for them. LO_ICP_NN27: get_closest_neighbour returns the nearest stored point over all 27 cells around the query's voxel
 * (cells visited x-outermost, points in storage order, first minimum wins). Every function below that looks neighbours up
 * (closest, correspondences, icp, the KissICP pipeline) follows the map's mode. */
enum { LO_ICP_REFERENCE = 0, LO_ICP_NN27 = 1 };
void lo_map_set_mode(lo_map *m, int icp_mode);
/* The plane of LO_ICP_PLANE = 2 (point-to-plane residual): normal of c >= 5 points (array-of-structs) = eigenvector of the smallest eigenvalue
 * of their scatter matrix by 5 cyclic Jacobi sweeps; returns 1 when planar (l_min <= 0.04 l_mid), 0 otherwise (nrm untouched). */
int lo_plane_normal(const double *pts, int c, double *nrm);
void lo_map_insert(lo_map *m, const double *xyz, long n);                    /* insert_points :12-62 */
void lo_map_update(lo_map *m, const double *xyz, long n, const double *pose7); /* update :138-144 */
void lo_map_remove_far(lo_map *m, const double *origin3);                    /* :146-171 under null locks */
/* get_closest_neighbour :64-102. out_xyz[3*n]; optional out_key[3*n] (voxel of the match, or
 * INT_MIN x3 when nothing was found and (0,0,0) is returned) and out_rank[n] (index of the match
 * inside its voxel, -1 when none). */
void lo_map_closest(const lo_map *m, const double *xyz, long n, double *out_xyz, int *out_key, int *out_rank);
/* get_correspondences :104-130 (index order). Optional out_idx[n]: source index of each pair. */
long lo_map_correspondences(const lo_map *m, const double *xyz, long n, double tau, double *src, double *tgt, long *out_idx);
long lo_map_dump(const lo_map *m, int *keys, int *counts, double *pts, long max_vox, long max_pts, long *n_pts);

/* ---- registration.cpp --------------------------------------------------------------------------- */
/* align_clouds :43-92. Any of H36 (row-major 6x6), g6, x6, pose7 may be NULL. */
void lo_align(const double *src, const double *tgt, long n, double th, double *H36, double *g6, double *x6, double *pose7);
/* ICP :94-130. Traces are optional: est_trace[7*max_iter], ncorr_trace[max_iter], hg_trace[42*max_iter]
 * (36 H row-major + 6 g per iteration), src_after[3*n]. Returns iterations executed. */
int lo_icp(const lo_map *m, const double *xyz, long n, const double *init7, double tau, double th, int max_iter,
           double eps, double *pose7, double *est_trace, long *ncorr_trace, double *hg_trace, double *src_after);

/* ---- deskew.cpp :10-28 ---------------------------------------------------------------------------- */
void lo_deskew(const float *xyz, const double *ts, long n, const double *T0, const double *T1, double *out);

/* ---- IMU-propagated backward deskew: per-point loop of kalman::EKF::motion_compensation_with_imu, L/src/kalman/ekf.cpp:420-468
 * (dead code at runtime in the reference). table: M rows of 22 doubles = kalman::Pose6D
 * {offset_time, acc[3], gyr[3], vel[3], pos[3], rot[9] row-major} (ekf.hpp:88-105) built by the host IMU forward pass (:315-391);
 * rot_end / pos_lidar_end: scan-end IMU rotation and lidar position (:393-418); p_il: IMU->lidar translation (state
 * POS_IMU_LIDAR). curv_ms: per-point offset time in ms (PCL `curvature`), points sorted by it (lidar/frame.cpp:28-51).
 * Writes the compensated xyz back as FLOAT (like :449-451) into xyz_f32 (in place) and widened to double into out. */
void lo_deskew_imu(float *xyz_f32, const float *curv_ms, long n, const double *table, long M, const double *rot_end9,
                   const double *pos_lidar_end3, const double *p_il3, double *out);

/* ---- frame::Lidar::process_frame, L/src/sensors/lidar/frame.cpp:101-193 (+ sort_clouds :28-51, split_clouds :53-99, utils::get_time_stamps
 * calculation_helpers.cpp:3-81): SURVEY section 8f N3, the host preprocessing in front of register_frame. limu_oracle_frame.c.
 * data: n point records `point_step` bytes apart; fields: nf PointField entries (names = NUL-terminated strings back to back).
 * cfg = {min_range, max_range, min_angle, max_angle, frame_rate, num_scan_lines, frame_split_num}; scan_count = the counter's value during
 * process_frame. Outputs for the concatenated segments: rec5[5*j] = {x, y, z, intensity, curvature}, ts[j]; per segment seg_sizes / seg_time.
 * Returns the number of segments, -1 when no timestamp field exists, -2 when a ring index exceeds num_scan_lines (constant-rotation path). */
long lo_process_frame(const unsigned char *data, long n, int point_step, int nf, const char *names, const int *offsets, const int *datatypes,
                      const int *counts, const double *cfg, double message_time, int scan_count, long max_segments, long *seg_sizes, double *seg_time,
                      float *rec5, double *ts);

/* ---- icp.cpp ---------------------------------------------------------------------------------------- */
long lo_voxel_downsample(const double *xyz, long n, double s, double *out, long *out_idx);   /* :9-30 */
long lo_iqr(const double *xyz, long n, double *out, double *bounds2);                         /* :88-124, common.hpp:22-63 */
void lo_voxelize(const double *xyz, long n, double v, double *src, long *n_src, double *down, long *n_down); /* :126-136 */

/* ---- threshold.cpp :5-28 ----------------------------------------------------------------------------- */
typedef struct lo_threshold { double init_threshold, min_motion_th, max_range, model_error_sq; int num_samples; double dev[7]; } lo_threshold;
void lo_threshold_init(lo_threshold *a, double init_th, double min_motion, double max_range);
double lo_threshold_step(lo_threshold *a, const double *dev7);   /* update_model_deviation + compute_threshold */

/* ---- KissICP pipeline: icp.cpp :36-86, :138-163 -------------------------------------------------------- */
typedef struct lo_kiss lo_kiss;
lo_kiss *lo_kiss_create(double voxel_size, double max_range, int cap, int deskew, double min_motion_th,
                        int icp_max_iteration, double initial_threshold, double estimation_threshold);
void lo_kiss_destroy(lo_kiss *k);
void lo_kiss_register_points(lo_kiss *k, const double *xyz, long n, double *down, long *n_down, double *src,
                             long *n_src, double *pose7);
void lo_kiss_register_cloud(lo_kiss *k, const float *xyz, const double *ts, long n, double *down, long *n_down,
                            double *src, long *n_src, double *pose7);
long lo_kiss_num_poses(const lo_kiss *k);
void lo_kiss_pose(const lo_kiss *k, long i, double *pose7);
lo_map *lo_kiss_map(lo_kiss *k);
void lo_kiss_set_mode(lo_kiss *k, int icp_mode);   /* LO_ICP_* of the local map */
int lo_kiss_last_iterations(const lo_kiss *k);
double lo_kiss_last_sigma(const lo_kiss *k);

#ifdef __cplusplus
}
#endif
#endif
