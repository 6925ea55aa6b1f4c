// C++ consumer of the drop-in kalman::EKF header (include/limu_dropin/limu/kalman/ekf.hpp) -- host code, runs without a GPU.
// Prints the state after a fixed script; tests/test_dropin_cpp.py compares it with the same script through the Python binding.
#include <cstdio>

#include "limu/kalman/ekf.hpp"

int main() {
    auto p = std::make_shared<kalman::EKF_PARAMETERS>();
    p->lidar_pose_trail = 3; p->noise_scale = 1.0;
    p->init_pos_noise = p->init_vel_noise = p->init_ori_noise = p->init_bga_noise = p->init_baa_noise = p->init_bat_noise = 1e-3;
    p->acc_process_noise = 0.03; p->gyro_process_noise = 0.00017; p->acc_process_noise_rev = 0.03; p->gyro_process_noise_rev = 0.00017;
    p->init_lidar_imu_time_noise = 1e-3; p->init_pos_trail_noise = 1e-3; p->init_ori_trail_noise = 1e-3; p->visualZuptR = 1e-3;
    kalman::EKF ekf(p);
    const Eigen::Vector3d grav(0, 0, -9.81), trans(0.05, -0.02, 0.1);
    Eigen::Matrix3d rot;
    rot << 0, -1, 0, 1, 0, 0, 0, 0, 1;
    ekf.initialize_imu_global_orientation(Eigen::Vector3d(0.3, -0.2, 9.7), grav);
    for (int i = 0; i < 20; ++i)
        ekf.predict(100.0 + 0.005 * i, Eigen::Vector3d(0.02, -0.01, 0.2 + 0.001 * i), Eigen::Vector3d(0.3, 0.1 * i, 9.81), grav, trans, rot);
    ekf.normalize_quaternions(true);
    ekf.update_and_propagate();
    const double pose[7] = {0.0, 0.0, 0.1, 0.99498743710662, 0.4, -0.1, 0.05};
    ekf.update_with_lidar_pose(pose, 0.05, 0.01);
    const Eigen::VectorXd m = ekf.state();
    const Eigen::MatrixXd P = ekf.covariance();
    for (int i = 0; i < m.size(); ++i) std::printf("%.17g\n", m[i]);
    std::printf("%.17g\n%.17g\n%.17g\n", P.trace(), ekf.get_current_time(), ekf.speed());
    return 0;
}
