// Drop-in replacement for L/include/limu/kalman/ekf.hpp (kalman::EKF): the public surface of the reference's filter -- constructor from
// EKF_PARAMETERS, predict, normalize_quaternions, update_and_propagate, the state accessors -- on top of liblimu_cuda's C ABI
// (limu_ekf_*, host code restated from L/src/kalman/ekf.cpp and pinned against the compiled original). The IMU-propagated deskew of the same
// class (motion_compensation_with_imu) is limu_imu_forward_pass + limu_deskew_imu.
// NEW, not in the reference: update_with_lidar_pose(), the registration result as a measurement (what the reference's design document promised).
#ifndef EKF_HPP
#define EKF_HPP

#include <Eigen/Dense>
#include <memory>
#include <stdexcept>
#include <string>

#include <limu_cuda.h>

namespace kalman
{
    constexpr int POS = 0, VEL = 3, ORI = 6, BGA = 10, BAA = 13, BAT = 16, GRAV = 19, POS_IMU_LIDAR = 22, ROT_IMU_LIDAR = 25, SFT = 29, LIDAR = 30;
    constexpr int INNER_DIM = LIDAR, POSE_DIM = 7;

    struct EKF_PARAMETERS   // ekf.hpp:62-86
    {
        typedef std::shared_ptr<EKF_PARAMETERS> Ptr;
        int lidar_pose_trail;
        double noise_scale;
        double init_pos_noise, init_vel_noise, init_ori_noise, init_bga_noise, init_baa_noise, init_bat_noise;
        double acc_process_noise, gyro_process_noise, acc_process_noise_rev, gyro_process_noise_rev;
        double init_lidar_imu_time_noise;
        double init_pos_trail_noise, init_ori_trail_noise;
        double visualZuptR;
    };

    class EKF
    {
    public:
        typedef std::unique_ptr<EKF> Ptr;

        explicit EKF(EKF_PARAMETERS::Ptr p)
        {
            limu_ekf_params q;
            limu_ekf_default_params(&q);
            q.lidar_pose_trail = p->lidar_pose_trail; q.noise_scale = p->noise_scale;
            q.init_pos_noise = p->init_pos_noise; q.init_vel_noise = p->init_vel_noise; q.init_ori_noise = p->init_ori_noise;
            q.init_bga_noise = p->init_bga_noise; q.init_baa_noise = p->init_baa_noise; q.init_bat_noise = p->init_bat_noise;
            q.acc_process_noise = p->acc_process_noise; q.gyro_process_noise = p->gyro_process_noise;
            q.acc_process_noise_rev = p->acc_process_noise_rev; q.gyro_process_noise_rev = p->gyro_process_noise_rev;
            q.init_lidar_imu_time_noise = p->init_lidar_imu_time_noise;
            q.init_pos_trail_noise = p->init_pos_trail_noise; q.init_ori_trail_noise = p->init_ori_trail_noise; q.visualZuptR = p->visualZuptR;
            ok(limu_ekf_create(&q, &h), "EKF");
            int32_t d = 0;
            ok(limu_ekf_state_dim(h, &d), "EKF");
            dim = d;
        }
        ~EKF() { limu_ekf_destroy(h); }
        EKF(const EKF &) = delete;
        EKF &operator=(const EKF &) = delete;

        void initialize_imu_global_orientation(const Eigen::Vector3d &xa, const Eigen::Vector3d &calc_grav)
        {
            ok(limu_ekf_initialize_orientation(h, xa.data(), calc_grav.data()), "initialize_imu_global_orientation");
        }
        void predict(double t, const Eigen::Vector3d &xg, const Eigen::Vector3d &xa, const Eigen::Vector3d &calc_grav,
                     const Eigen::Vector3d &trans_lidar_imu, const Eigen::Matrix3d &rot_lidar_imu)
        {
            const Eigen::Matrix<double, 3, 3, Eigen::RowMajor> R = rot_lidar_imu;
            ok(limu_ekf_predict(h, t, xg.data(), xa.data(), calc_grav.data(), trans_lidar_imu.data(), R.data()), "predict");
        }
        void normalize_quaternions(bool only_current = false) { ok(limu_ekf_normalize_quaternions(h, only_current ? 1 : 0), "normalize_quaternions"); }
        void update_and_propagate() { ok(limu_ekf_update_and_propagate(h), "update_and_propagate"); }
        // not in the reference: pose = {qx,qy,qz,qw, tx,ty,tz} from lidar::KissICP::register_frame as a measurement of POS / ORI
        void update_with_lidar_pose(const double pose[7], double pos_sigma, double ori_sigma) { ok(limu_ekf_update_lidar_pose(h, pose, pos_sigma, ori_sigma), "update_with_lidar_pose"); }

        Eigen::VectorXd state() const
        {
            Eigen::VectorXd m(dim);
            ok(limu_ekf_get_state(h, m.data(), nullptr, nullptr), "state");
            return m;
        }
        Eigen::MatrixXd covariance() const
        {
            Eigen::Matrix<double, Eigen::Dynamic, Eigen::Dynamic, Eigen::RowMajor> P(dim, dim);
            ok(limu_ekf_get_state(h, nullptr, P.data(), nullptr), "covariance");
            return P;
        }
        Eigen::Vector3d position() const { return state().segment(POS, 3); }
        Eigen::Vector3d velocity() const { return state().segment(VEL, 3); }
        Eigen::Vector4d orientation() const { return state().segment(ORI, 4); }
        Eigen::Vector3d gravity_check() const { return state().segment(GRAV, 3); }
        double speed() const { return velocity().norm(); }
        double get_current_time() const
        {
            double t = 0.0;
            ok(limu_ekf_get_state(h, nullptr, nullptr, &t), "get_current_time");
            return t;
        }
        limu_ekf *handle() const { return h; }

    private:
        static void ok(int st, const char *what)
        {
            if (st != LIMU_OK) throw std::runtime_error(std::string("kalman::EKF::") + what + ": " + limu_last_error());
        }
        limu_ekf *h = nullptr;
        int dim = 0;
    };
}
#endif
