// imu_forward.cu -- HOST code: the IMU forward pass of kalman::EKF::motion_compensation_with_imu (L/src/kalman/ekf.cpp:292-418),
// i.e. what produces the pose table, the scan-end rotation and the scan-end lidar position that the per-point kernel
// (imu_deskew.cu, ekf.cpp:420-468) consumes. SURVEY section 8f N1. It walks ~20 IMU samples per scan: scalar work that stays on the
// host (like the reference), but inside the library, so a caller no longer needs the reference's EKF object to deskew a scan with IMU data.
//
// Restated, statement by statement:
//   * per IMU pair (head, tail) with tail newer than the previous scan's end (:316-326): mid-point rates (:328-335), dt (:340-343),
//     accelerations rescaled by gravity / |mean_acc| (:357), quaternion update prev_quat <- exp(S(-dt)) prev_quat (:373-375),
//     rot = Quaterniond(prev_quat).toRotationMatrix() (:376, the 4-vector is taken as x,y,z,w and NOT normalised, helper.hpp:11-17),
//     vel += (rot^T T_ab + grav) dt, pos += vel dt (:379-383), one kalman::Pose6D row (:390);
//   * scan end (:393-410): dt = |pcl_end - imu_end|, one more quaternion update with S(+dt) -- the sign differs from the loop's, as in the
//     reference --, vel_end, pos_end, pos_lidar_end = rot_end p_IL + pos_end (:417).
// S(dt) = -(dt/2) Omega(w - b_g) with Omega 4x4 skew and Omega^2 = -|w|^2 I (ekf.cpp:471-484), so exp(S) = cos(theta) I + sin(theta)/theta S
// with theta = |dt| |w| / 2 in closed form; the reference evaluates the same exponential with Eigen's Pade approximant (unsupported
// MatrixFunctions), which agrees to ~1e-16. The covariance the reference propagates alongside (:345-370) goes into a local copy that is
// never written back and is not reproduced.
#include <math.h>

#include "common.cuh"

namespace {

void quat_to_rot(const double *q /* x y z w as Eigen reads the 4-vector */, double *R /* row-major */) {
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
    R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}

// prev_quat <- exp(calculate_S(xg, m, dt)) * prev_quat
void quat_step(const double *xg, const double *bga, double dt, double *q) {
    const double w[3] = {xg[0] - bga[0], xg[1] - bga[1], xg[2] - bga[2]};
    const double c = -dt / 2;
    const double S[16] = {0, -w[0] * c, -w[1] * c, -w[2] * c, w[0] * c, 0, -w[2] * c, w[1] * c,
                          w[1] * c, w[2] * c, 0, -w[0] * c, w[2] * c, -w[1] * c, w[0] * c, 0};
    const double theta = fabs(c) * sqrt((w[0] * w[0] + w[1] * w[1]) + w[2] * w[2]);
    const double cs = cos(theta), sc = theta > 0.0 ? sin(theta) / theta : 1.0;
    double out[4];
    for (int r = 0; r < 4; ++r) {
        double acc = cs * q[r];
        for (int k = 0; k < 4; ++k) acc += sc * S[4 * r + k] * q[k];
        out[r] = acc;
    }
    for (int r = 0; r < 4; ++r) q[r] = out[r];
}

void add_row(limu_imu_pose *row, double offset, const double *acc, const double *gyr, const double *vel, const double *pos, const double *R) {
    row->offset_time = offset;
    for (int a = 0; a < 3; ++a) { row->acc[a] = acc[a]; row->gyr[a] = gyr[a]; row->vel[a] = vel[a]; row->pos[a] = pos[a]; }
    for (int k = 0; k < 9; ++k) row->rot[k] = R[k];
}

}  // namespace

extern "C" int limu_imu_forward_pass(limu_imu_state *st, const limu_imu_sample *s, int32_t k, double lidar_beg_time, double last_point_curvature_ms,
                                     limu_imu_pose *table, int32_t max_rows, int32_t *n_rows, double rot_end[9], double pos_lidar_end[3]) {
    LIMU_REQUIRE(st && s && table && n_rows && rot_end && pos_lidar_end && k >= 2 && max_rows >= 2, "limu_imu_forward_pass: bad arguments (need k >= 2 samples, sample 0 = the last one of the previous window)");
    LIMU_REQUIRE(st->mean_acc_norm > 0.0, "limu_imu_forward_pass: mean_acc_norm must be positive");
    const double pcl_end_time = lidar_beg_time + last_point_curvature_ms / double(1000);   // :299
    const double imu_end_time = s[k - 1].t;
    double vel[3], pos[3], q[4], R[9];
    for (int a = 0; a < 3; ++a) { vel[a] = st->vel[a]; pos[a] = st->pos[a]; }
    for (int a = 0; a < 4; ++a) q[a] = st->quat[a];
    quat_to_rot(q, R);                                                                     // :305
    int rows = 0;
    add_row(&table[rows++], 0.0, st->acc_s_last, st->ang_vel_last, vel, pos, R);           // populate_imu_pose(0.0) :307
    double xa[3] = {0, 0, 0}, xg[3] = {0, 0, 0}, dt = 0.0;
    for (int r = 0; r + 1 < k; ++r) {
        const limu_imu_sample &head = s[r], &tail = s[r + 1];
        if (tail.t < st->last_lidar_end_time) continue;                                    // :322-323
        for (int a = 0; a < 3; ++a) { xg[a] = 0.5 * (head.gyr[a] + tail.gyr[a]); xa[a] = 0.5 * (head.acc[a] + tail.acc[a]); }
        dt = head.t < st->last_lidar_end_time ? tail.t - st->last_lidar_end_time : tail.t - head.t;   // :340-343
        for (int a = 0; a < 3; ++a) xa[a] = xa[a] / st->mean_acc_norm * st->gravity_norm;        // :357
        quat_step(xg, st->bga, -dt, q);                                                    // :373-375
        quat_to_rot(q, R);                                                                 // :376
        double T_ab[3];
        for (int a = 0; a < 3; ++a) T_ab[a] = st->bat[a] * xa[a] - st->baa[a];             // :379
        for (int a = 0; a < 3; ++a) vel[a] += ((R[a] * T_ab[0] + R[3 + a] * T_ab[1] + R[6 + a] * T_ab[2]) + st->grav[a]) * dt;   // rot^T T_ab :380
        for (int a = 0; a < 3; ++a) pos[a] += vel[a] * dt;                                 // :383
        for (int a = 0; a < 3; ++a) { st->acc_s_last[a] = xa[a]; st->ang_vel_last[a] = xg[a]; }   // :386-387
        if (rows >= max_rows) { limu::set_error("limu_imu_forward_pass: the pose table needs more than %d rows", max_rows); return LIMU_ERR_INVALID; }
        add_row(&table[rows++], tail.t - lidar_beg_time, xa, xg, vel, pos, R);              // :390
    }
    // position and attitude at the frame end (:393-410)
    dt = (pcl_end_time > imu_end_time ? 1.0 : -1.0) * (pcl_end_time - imu_end_time);
    quat_step(xg, st->bga, dt, q);
    quat_to_rot(q, rot_end);
    double T_ab[3], vel_end[3], pos_end[3];
    for (int a = 0; a < 3; ++a) T_ab[a] = st->bat[a] * xa[a] - st->baa[a];
    for (int a = 0; a < 3; ++a) vel_end[a] = vel[a] + ((rot_end[a] * T_ab[0] + rot_end[3 + a] * T_ab[1] + rot_end[6 + a] * T_ab[2]) + st->grav[a]) * dt;
    for (int a = 0; a < 3; ++a) pos_end[a] = pos[a] + vel_end[a] * dt;
    for (int a = 0; a < 3; ++a)
        pos_lidar_end[a] = (rot_end[3 * a] * st->p_imu_lidar[0] + rot_end[3 * a + 1] * st->p_imu_lidar[1] + rot_end[3 * a + 2] * st->p_imu_lidar[2]) + pos_end[a];   // :417
    // what the reference carries to the next window (:412-414 and the tracker members)
    st->last_lidar_end_time = pcl_end_time;
    for (int a = 0; a < 3; ++a) { st->tracker_vel[a] = vel[a]; st->tracker_pos[a] = pos[a]; }
    for (int a = 0; a < 4; ++a) st->tracker_quat[a] = q[a];
    *n_rows = rows;
    return LIMU_OK;
}
