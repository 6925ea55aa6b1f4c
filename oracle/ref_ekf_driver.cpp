// Oracle (test infrastructure): C entry point over the reference's UNMODIFIED kalman::EKF::motion_compensation_with_imu
// (L/src/kalman/ekf.cpp:292-469) -- the IMU-propagated backward deskew that SURVEY section 8(f) lists as N1. The function is
// dead at runtime in the reference (its only caller is never invoked) but it compiles and runs; this driver builds the
// EKF with its constructor defaults, feeds it one scan + IMU window, and returns
//   * the deskewed points the reference produced (meas->deskewed),
//   * the IMU pose table it built (mc_tracker->imu_pose, a public member),
//   * the scan-end rotation / lidar position, which are LOCALS of the reference function: they are recomputed here with
//     the reference's own calculate_S + Eigen's matrix exponential, statement for statement (ekf.cpp:336-337,373-375,
//     393-418). `#define private public` only lifts access control; no reference source is modified or copied.
// everything ekf.hpp pulls in is included FIRST (include guards), so the access lift below touches only ekf.hpp's own text
#include <array>
#include <deque>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>
#include <eigen3/unsupported/Eigen/MatrixFunctions>
#include <Eigen/SparseCore>
#include <Eigen/Cholesky>
#include <sensor_msgs/Imu.h>
#include "common.hpp"
#include "limu/sensors/sync_frame.hpp"
#define private public
#include "limu/kalman/ekf.hpp"
#undef private
#include "limu/kalman/helper.hpp"

#include <cstring>

static double g_last_lidar_end_time = 0.0;

extern "C" {

// EKF::last_lidar_end_time as the NEXT ref_imu_deskew call finds it (what the previous window's call left behind, ekf.cpp:413): lets a
// test reach the "pair older than the previous scan's end" branches (:322-323, :340-341) without a long-lived EKF object.
void ref_imu_set_last_lidar_end_time(double t) { g_last_lidar_end_time = t; }

// imu: k rows of {t, gx, gy, gz, ax, ay, az}; row 0 plays mc_tracker->last_imu (the sample before the window).
// xyz/curv_ms: n points sorted by curvature (per-point offset time in ms, lidar/frame.cpp:28-51).
// table_out: M rows of 22 doubles {offset_time, acc3, gyr3, vel3, pos3, rot9 row-major}; returns M.
long ref_imu_deskew(const float *xyz, const float *curv_ms, long n, const double *imu, long k, double lidar_beg_time, const double *mean_acc3,
                    const double *p_imu_lidar3, const double *gyro_bias3, double *deskewed_out, double *table_out, long max_rows, double *rot_end9,
                    double *pos_lidar_end3, float *xyz_written_back) {
    auto prm = std::make_shared<kalman::EKF_PARAMETERS>();
    std::memset(prm.get(), 0, sizeof(kalman::EKF_PARAMETERS));
    prm->lidar_pose_trail = 20;   // odom_run.cpp:19 default
    prm->noise_scale = 1.0;
    prm->init_pos_noise = prm->init_vel_noise = prm->init_ori_noise = prm->init_bga_noise = prm->init_baa_noise = prm->init_bat_noise = 1e-3;
    prm->acc_process_noise = 0.03; prm->gyro_process_noise = 0.00017;   // odom_run.cpp:27-28
    prm->acc_process_noise_rev = 0.03; prm->gyro_process_noise_rev = 0.00017;
    prm->init_lidar_imu_time_noise = 1e-3; prm->init_pos_trail_noise = 1e-3; prm->init_ori_trail_noise = 1e-3; prm->visualZuptR = 1e-3;
    kalman::EKF ekf(prm);
    for (int i = 0; i < 3; ++i) { ekf.m(kalman::POS_IMU_LIDAR + i) = p_imu_lidar3[i]; ekf.m(kalman::BGA + i) = gyro_bias3[i]; }
    ekf.m.segment(kalman::GRAV, 3) = ekf.grav;   // gravity in the state (what initialize_imu_global_orientation would have set)
    ekf.last_lidar_end_time = g_last_lidar_end_time;

    auto mk = [&](long r) {
        auto p = std::make_shared<sensor_msgs::Imu>();
        p->header.stamp.fromSec(imu[7 * r]);
        p->angular_velocity.x = imu[7 * r + 1]; p->angular_velocity.y = imu[7 * r + 2]; p->angular_velocity.z = imu[7 * r + 3];
        p->linear_acceleration.x = imu[7 * r + 4]; p->linear_acceleration.y = imu[7 * r + 5]; p->linear_acceleration.z = imu[7 * r + 6];
        return p;
    };
    ekf.mc_tracker->last_imu = mk(0);
    frame::LidarImuInit::Ptr meas(new frame::LidarImuInit());
    meas->lidar_beg_time = lidar_beg_time;
    meas->mean_acc = utils::Vec3d(mean_acc3[0], mean_acc3[1], mean_acc3[2]);
    for (long r = 1; r < k; ++r) meas->imu_buffer.push_back(mk(r));
    meas->processed_frame->points.resize(static_cast<size_t>(n));
    for (long i = 0; i < n; ++i) {
        auto &p = meas->processed_frame->points[static_cast<size_t>(i)];
        p.x = xyz[3 * i]; p.y = xyz[3 * i + 1]; p.z = xyz[3 * i + 2]; p.curvature = curv_ms[i];
    }

    // ---- scan-end state, recomputed with the reference's own pieces BEFORE the call mutates anything we read ----
    {
        const double last_lidar_end_time = ekf.last_lidar_end_time;
        const double pcl_end_time = lidar_beg_time + double(curv_ms[n - 1]) / double(1000);
        const double imu_end_time = imu[7 * (k - 1)];
        Eigen::Vector4d prev_quat = ekf.orientation();
        Eigen::Vector3d vel = ekf.velocity(), pos = ekf.position(), xg = Eigen::Vector3d::Zero(), xa = Eigen::Vector3d::Zero();
        Eigen::Matrix3d rot;
        double dt = 0;
        for (long r = 0; r + 1 < k; ++r) {   // v_imu = {last_imu, buffer...}; pairs (head, tail)
            const double head_t = imu[7 * r], tail_t = imu[7 * (r + 1)];
            if (tail_t < last_lidar_end_time) continue;
            for (int a = 0; a < 3; ++a) { xg[a] = 0.5 * (imu[7 * r + 1 + a] + imu[7 * (r + 1) + 1 + a]); xa[a] = 0.5 * (imu[7 * r + 4 + a] + imu[7 * (r + 1) + 4 + a]); }
            dt = head_t < last_lidar_end_time ? tail_t - last_lidar_end_time : tail_t - head_t;
            xa = xa / meas->get_mean_acc_norm() * gravity;
            Eigen::Matrix4d S = ekf.calculate_S(xg, ekf.m, -dt);
            Eigen::Matrix4d A = S.exp();
            prev_quat = A * prev_quat;
            rot = utils::quat2rmat(prev_quat);
            Eigen::Vector3d T_ab = ekf.m.segment(kalman::BAT, 3).asDiagonal() * xa - ekf.m.segment(kalman::BAA, 3);
            vel += (rot.transpose() * T_ab + ekf.m.segment(kalman::GRAV, 3)) * dt;
            pos += vel * dt;
        }
        const double note = pcl_end_time > imu_end_time ? 1.0 : -1.0;
        dt = note * (pcl_end_time - imu_end_time);
        Eigen::Matrix4d S = ekf.calculate_S(xg, ekf.m, dt);
        Eigen::Matrix4d A = S.exp();
        prev_quat = A * prev_quat;
        const Eigen::Matrix3d rot_end = utils::quat2rmat(prev_quat);
        Eigen::Vector3d T_ab = ekf.m.segment(kalman::BAT, 3).asDiagonal() * xa - ekf.m.segment(kalman::BAA, 3);
        const Eigen::Vector3d vel_end = vel + (rot_end.transpose() * T_ab + ekf.m.segment(kalman::GRAV, 3)) * dt;
        const Eigen::Vector3d pos_end = pos + vel_end * dt;
        const Eigen::Vector3d pos_lidar_end = rot_end * ekf.m.segment(kalman::POS_IMU_LIDAR, 3) + pos_end;
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) rot_end9[3 * r + c] = rot_end(r, c);
        for (int a = 0; a < 3; ++a) pos_lidar_end3[a] = pos_lidar_end[a];
    }

    ekf.motion_compensation_with_imu(meas);   // the reference

    for (long i = 0; i < n; ++i) {
        const auto &d = meas->deskewed[static_cast<size_t>(i)];
        deskewed_out[3 * i] = d[0]; deskewed_out[3 * i + 1] = d[1]; deskewed_out[3 * i + 2] = d[2];
        if (xyz_written_back) {
            const auto &p = meas->processed_frame->points[static_cast<size_t>(i)];
            xyz_written_back[3 * i] = p.x; xyz_written_back[3 * i + 1] = p.y; xyz_written_back[3 * i + 2] = p.z;
        }
    }
    const auto &tab = ekf.mc_tracker->imu_pose;
    const long M = static_cast<long>(tab.size());
    for (long r = 0; r < M && r < max_rows; ++r) {
        double *o = table_out + 22 * r;
        o[0] = tab[r].offset_time;
        for (int a = 0; a < 3; ++a) { o[1 + a] = tab[r].acc[a]; o[4 + a] = tab[r].gyr[a]; o[7 + a] = tab[r].vel[a]; o[10 + a] = tab[r].pos[a]; }
        for (int rr = 0; rr < 3; ++rr) for (int c = 0; c < 3; ++c) o[13 + 3 * rr + c] = tab[r].rot(rr, c);
    }
    return M;
}

// ---- kalman::EKF predict / update: a long-lived reference object behind a C handle -------------------------------
// params: {lidar_pose_trail, noise_scale, init_pos, init_vel, init_ori, init_bga, init_baa, init_bat, acc_noise, gyro_noise, acc_rev, gyro_rev,
//          init_lidar_imu_time, init_pos_trail, init_ori_trail, visualZuptR} as 16 doubles.
void *ref_ekf_create(const double *prm16) {
    auto prm = std::make_shared<kalman::EKF_PARAMETERS>();
    std::memset(prm.get(), 0, sizeof(kalman::EKF_PARAMETERS));
    prm->lidar_pose_trail = static_cast<int>(prm16[0]);
    prm->noise_scale = prm16[1];
    prm->init_pos_noise = prm16[2]; prm->init_vel_noise = prm16[3]; prm->init_ori_noise = prm16[4];
    prm->init_bga_noise = prm16[5]; prm->init_baa_noise = prm16[6]; prm->init_bat_noise = prm16[7];
    prm->acc_process_noise = prm16[8]; prm->gyro_process_noise = prm16[9]; prm->acc_process_noise_rev = prm16[10]; prm->gyro_process_noise_rev = prm16[11];
    prm->init_lidar_imu_time_noise = prm16[12]; prm->init_pos_trail_noise = prm16[13]; prm->init_ori_trail_noise = prm16[14]; prm->visualZuptR = prm16[15];
    return new kalman::EKF(prm);
}
void ref_ekf_destroy(void *h) { delete static_cast<kalman::EKF *>(h); }
long ref_ekf_dim(void *h) { return static_cast<kalman::EKF *>(h)->state_dim; }
void ref_ekf_get(void *h, double *m, double *P /* row-major */, double *time) {
    auto &e = *static_cast<kalman::EKF *>(h);
    const int n = e.state_dim;
    for (int i = 0; i < n; ++i) m[i] = e.m(i);
    for (int r = 0; r < n; ++r) for (int c = 0; c < n; ++c) P[(size_t)r * n + c] = e.P(r, c);
    if (time) *time = e.get_current_time();
}
void ref_ekf_set(void *h, const double *m, const double *P) {
    auto &e = *static_cast<kalman::EKF *>(h);
    const int n = e.state_dim;
    if (m) for (int i = 0; i < n; ++i) e.m(i) = m[i];
    if (P) for (int r = 0; r < n; ++r) for (int c = 0; c < n; ++c) e.P(r, c) = P[(size_t)r * n + c];
}
void ref_ekf_predict(void *h, double t, const double *xg, const double *xa, const double *grav, const double *trans, const double *rot9 /* row-major */) {
    auto &e = *static_cast<kalman::EKF *>(h);
    Eigen::Matrix3d R;
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) R(r, c) = rot9[3 * r + c];
    e.predict(t, Eigen::Vector3d(xg[0], xg[1], xg[2]), Eigen::Vector3d(xa[0], xa[1], xa[2]), Eigen::Vector3d(grav[0], grav[1], grav[2]),
              Eigen::Vector3d(trans[0], trans[1], trans[2]), R);
}
void ref_ekf_normalize(void *h, int only_current) { static_cast<kalman::EKF *>(h)->normalize_quaternions(only_current != 0); }
void ref_ekf_zero_vel_update(void *h, double r) {
    auto &e = *static_cast<kalman::EKF *>(h);
    e.zero_vel_update(e.m, e.K, e.P, e.HP, e.invS, e.H, e.R, r);
}
void ref_ekf_augment(void *h) { static_cast<kalman::EKF *>(h)->update_visual_pose_aug(); }
// update_undo_augmentation pops augment_times unconditionally (undefined on the empty vector the reference's own bookkeeping leaves): give it an entry to pop
void ref_ekf_undo_augmentation(void *h) {
    auto &e = *static_cast<kalman::EKF *>(h);
    if (e.augment_times.empty()) { e.augment_times.push_back(0.0); e.augment_count++; }
    e.update_undo_augmentation();
}

}  // extern "C"
