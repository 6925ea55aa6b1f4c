// ekf.cu -- HOST code: kalman::EKF predict / zero-velocity update / pose-trail augmentation (L/src/kalman/ekf.cpp:63-290, 471-764), SURVEY
// section 8f N4, restated without Eigen. The north star keeps "the small state solve" on the host: this is 30 + 7 * trail scalars (170 with
// the reference's default trail of 20) and dense products of that size, once per IMU sample / per scan. In the reference this filter is
// constructed but never stepped at runtime; it compiles, and oracle/ref_ekf_driver.cpp drives the compiled original for
// parity. What the reference's design promised and never wired -- the registration result entering the filter -- is limu_ekf_update_lidar_pose
// (no counterpart: defined here, parity unpinned).
//
// State layout (ekf.hpp:32-44): POS 0, VEL 3, ORI 6 (w x y z), BGA 10, BAA 13, BAT 16, GRAV 19, POS_IMU_LIDAR 22, ROT_IMU_LIDAR 25, SFT 29,
// then `trail` poses of 7 (position 3 + orientation 4) from 30. Matrices are row-major.
//
// Faithful quirks (both arms alike): the quaternion 4-vector of the state is handed to Eigen::Quaterniond(Vector4d), which reads it as
// x,y,z,w and does not normalise (helper.hpp:11-17); the "derivative" dR[i] is quat2rmat(e_i) - quat2rmat(q) (helper.hpp:19-33); BGA
// mean-reverts with gyro_process_noise, not its _rev twin (ekf.cpp:511-512); row/column SFT of dydx stays zero. exp(S) is closed-form here
// (S^2 = -theta^2 I), Eigen's Pade approximant there: ~1e-16 apart.
#include <math.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace {

enum { POS = 0, VEL = 3, ORI = 6, BGA = 10, BAA = 13, BAT = 16, GRAV = 19, POS_IMU_LIDAR = 22, ROT_IMU_LIDAR = 25, SFT = 29, LIDAR = 30, INNER = 30, POSE_DIM = 7 };
enum { Q_ACC = 0, Q_GYRO = 3, Q_BGA_DRIFT = 6, Q_BAA_DRIFT = 9, Q_DIM = 12 };

inline double sq(double x) { return x * x; }

void quat_to_rot(const double *q /* x y z w as Eigen reads the 4-vector */, double *R /* row-major */) {
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
    R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}

// Eigen::Quaterniond(Matrix3d) (Quaternion.h, quaternionbase_assign_impl<Matrix3>): coefficients x, y, z, w
void rot_to_quat(const double *m /* row-major */, double *c /* x y z w */) {
    double t = (m[0] + m[4]) + m[8];
    if (t > 0) {
        t = sqrt(t + 1.0);
        c[3] = 0.5 * t; t = 0.5 / t;
        c[0] = (m[7] - m[5]) * t; c[1] = (m[2] - m[6]) * t; c[2] = (m[3] - m[1]) * t;
    } else {
        int i = 0;
        if (m[4] > m[0]) i = 1;
        if (m[8] > m[4 * i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = sqrt(m[4 * i] - m[4 * j] - m[4 * k] + 1.0);
        c[i] = 0.5 * t; t = 0.5 / t;
        c[3] = (m[3 * k + j] - m[3 * j + k]) * t; c[j] = (m[3 * j + i] + m[3 * i + j]) * t; c[k] = (m[3 * k + i] + m[3 * i + k]) * t;
    }
}

// S = -(dt/2) Omega(xg - bga) (calculate_S, ekf.cpp:471-484) and A = exp(S) in closed form
void calc_A(const double *xg, const double *bga, double dt, double *A /* 4x4 */) {
    const double w[3] = {xg[0] - bga[0], xg[1] - bga[1], xg[2] - bga[2]};
    const double c = -dt / 2;
    const double S[16] = {0, -w[0] * c, -w[1] * c, -w[2] * c, w[0] * c, 0, -w[2] * c, w[1] * c,
                          w[1] * c, w[2] * c, 0, -w[0] * c, w[2] * c, -w[1] * c, w[0] * c, 0};
    const double theta = fabs(c) * sqrt((w[0] * w[0] + w[1] * w[1]) + w[2] * w[2]);
    const double cs = cos(theta), sc = theta > 0.0 ? sin(theta) / theta : 1.0;
    for (int i = 0; i < 16; ++i) A[i] = sc * S[i] + ((i % 5 == 0) ? cs : 0.0);
}

// X = S^-1 B for a small symmetric S (n x n, row-major) and B (n x cols): Eigen 3.4.0 LDLT<MatrixXd>::compute + solve restated (Cholesky/LDLT.h:300-395,
// 569-607: in-place on the lower triangle, pivoting on the largest remaining diagonal entry, pseudo-inverse of D), which is what ekf.cpp:45-47
// and :717-718 call. With R = 1e-9 next to unit variances the innovation covariance has a condition number ~1e6: an unpivoted factorisation
// is just as valid but lands ~3e-10 away from Eigen's answer, the same pivot order lands within rounding.
bool ldlt_solve(int n, std::vector<double> a, const double *B, int cols, double *X) {
    auto A = [&](int r, int c) -> double & { return a[(size_t)r * n + c]; };
    std::vector<int> tr(n);
    std::vector<double> temp(n);
    for (int k = 0; k < n; ++k) {
        int big = k;
        double bigv = fabs(A(k, k));
        for (int i = k + 1; i < n; ++i) { const double v = fabs(A(i, i)); if (v > bigv) { bigv = v; big = i; } }
        tr[k] = big;
        if (k != big) {
            for (int c = 0; c < k; ++c) std::swap(A(k, c), A(big, c));
            for (int r = big + 1; r < n; ++r) std::swap(A(r, k), A(r, big));
            std::swap(A(k, k), A(big, big));
            for (int i = k + 1; i < big; ++i) std::swap(A(i, k), A(big, i));
        }
        if (k > 0) {
            for (int c = 0; c < k; ++c) temp[c] = A(c, c) * A(k, c);
            double acc = 0.0;
            for (int c = 0; c < k; ++c) acc += A(k, c) * temp[c];
            A(k, k) -= acc;
            for (int r = k + 1; r < n; ++r) {
                double a2 = 0.0;
                for (int c = 0; c < k; ++c) a2 += A(r, c) * temp[c];
                A(r, k) -= a2;
            }
        }
        const double akk = A(k, k);
        if (k == 0 && !(fabs(akk) > 0.0)) return false;   // entire diagonal zero
        if (fabs(akk) > 0.0) for (int r = k + 1; r < n; ++r) A(r, k) /= akk;
    }
    std::vector<double> d(n);
    for (int c = 0; c < cols; ++c) {
        for (int i = 0; i < n; ++i) d[i] = B[(size_t)i * cols + c];
        for (int k = 0; k < n; ++k) std::swap(d[k], d[tr[k]]);
        for (int i = 0; i < n; ++i) for (int q = 0; q < i; ++q) d[i] -= A(i, q) * d[q];
        for (int i = 0; i < n; ++i) d[i] = fabs(A(i, i)) > 2.2250738585072014e-308 ? d[i] / A(i, i) : 0.0;
        for (int i = n - 1; i >= 0; --i) for (int q = i + 1; q < n; ++q) d[i] -= A(q, i) * d[q];
        for (int k = n - 1; k >= 0; --k) std::swap(d[k], d[tr[k]]);
        for (int i = 0; i < n; ++i) X[(size_t)i * cols + c] = d[i];
    }
    return true;
}

void normalize4(double *q) {   // Eigen normalize(): untouched when the squared norm is not positive
    const double z = (q[0] * q[0] + q[1] * q[1]) + (q[2] * q[2] + q[3] * q[3]);
    if (z > 0) { const double n = sqrt(z); for (int i = 0; i < 4; ++i) q[i] /= n; }
}

}  // namespace

struct limu_ekf {
    limu_ekf_params prm;
    int trail = 0, dim = 0;
    double noise_scale = 1.0;   // params->noise_scale squared (ekf.cpp:65)
    std::vector<double> m, P, Q, dydx, dydq;
    double time = 0.0, ZUPTtime = -1.0, prev_sampleT = -1.0, first_sampleT = -1.0;
    bool first_sample = true, was_stationary = false;
    std::vector<double> augment_times;
    int augment_count = 0;
    double &Pm(int r, int c) { return P[(size_t)r * dim + c]; }
    void normalize_quaternions(bool only_current) {   // ekf.cpp:619-636
        normalize4(&m[ORI]);
        normalize4(&m[ROT_IMU_LIDAR]);
        if (only_current) return;
        for (int i = 0; i < trail; ++i) normalize4(&m[LIDAR + POSE_DIM * i + 3]);
    }
    void symmetrize() {   // maintain_positive_semi_definite, ekf.cpp:758-764
        for (int r = 0; r < dim; ++r)
            for (int c = r + 1; c < dim; ++c) { const double v = 0.5 * (Pm(r, c) + Pm(c, r)); Pm(r, c) = v; Pm(c, r) = v; }
    }
    // m <- A m, P <- A P A^T for a 0/1 matrix A with at most one entry per row: new index i takes old index src[i] (or nothing)
    void remap(const std::vector<int> &src) {
        std::vector<double> m2(dim, 0.0), P2((size_t)dim * dim, 0.0);
        for (int r = 0; r < dim; ++r) {
            if (src[r] < 0) continue;
            m2[r] = m[src[r]];
            for (int c = 0; c < dim; ++c)
                if (src[c] >= 0) P2[(size_t)r * dim + c] = P[(size_t)src[r] * dim + src[c]];
        }
        m.swap(m2);
        P.swap(P2);
    }
    // the anonymous update() of ekf.cpp:36-61 for a measurement matrix H (rows x l, l <= dim leading state entries), y (rows), R (rows x rows)
    int update(int rows, int l, const std::vector<double> &H, const std::vector<double> &y, const std::vector<double> &R, bool joseph) {
        std::vector<double> HP((size_t)rows * dim, 0.0), S(R), K((size_t)dim * rows), X((size_t)rows * dim);
        for (int r = 0; r < rows; ++r)
            for (int k = 0; k < l; ++k) {
                const double h = H[(size_t)r * l + k];
                if (h == 0.0) continue;
                for (int c = 0; c < dim; ++c) HP[(size_t)r * dim + c] += h * P[(size_t)k * dim + c];
            }
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < rows; ++c)
                for (int k = 0; k < l; ++k) {   // accumulated INTO R term by term, like Eigen's noalias() += with its sparse H: (R + a) - b, not R + (a - b)
                    const double h = H[(size_t)c * l + k];
                    if (h != 0.0) S[(size_t)r * rows + c] += HP[(size_t)r * dim + k] * h;
                }
        if (!ldlt_solve(rows, S, HP.data(), dim, X.data())) { limu::set_error("limu_ekf: innovation covariance is singular"); return LIMU_ERR_INVALID; }
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < dim; ++c) K[(size_t)c * rows + r] = X[(size_t)r * dim + c];   // K = (S^-1 HP)^T
        std::vector<double> v(rows);
        for (int r = 0; r < rows; ++r) {
            double a = 0.0;
            for (int k = 0; k < l; ++k) a += H[(size_t)r * l + k] * m[k];
            v[r] = y[r] - a;
        }
        for (int i = 0; i < dim; ++i) {
            double a = 0.0;
            for (int r = 0; r < rows; ++r) a += K[(size_t)i * rows + r] * v[r];
            m[i] += a;
        }
        if (!joseph) {   // update_common, ekf.cpp:10-17
            for (int i = 0; i < dim; ++i)
                for (int r = 0; r < rows; ++r) {
                    const double k = K[(size_t)i * rows + r];
                    if (k == 0.0) continue;
                    for (int c = 0; c < dim; ++c) P[(size_t)i * dim + c] -= k * HP[(size_t)r * dim + c];
                }
            normalize4(&m[ORI]);
            normalize4(&m[ROT_IMU_LIDAR]);
            return LIMU_OK;
        }
        // update_common_joseph_form, ekf.cpp:19-34: P = (I - K H) P (I - K H)^T + K R K^T
        std::vector<double> T((size_t)dim * dim, 0.0), TP((size_t)dim * dim, 0.0), P2((size_t)dim * dim, 0.0);
        for (int i = 0; i < dim; ++i) {
            for (int r = 0; r < rows; ++r) {
                const double k = K[(size_t)i * rows + r];
                for (int c = 0; c < l; ++c) T[(size_t)i * dim + c] -= k * H[(size_t)r * l + c];
            }
            T[(size_t)i * dim + i] += 1.0;
        }
        for (int i = 0; i < dim; ++i)
            for (int k = 0; k < dim; ++k) {
                const double t = T[(size_t)i * dim + k];
                if (t == 0.0) continue;
                for (int c = 0; c < dim; ++c) TP[(size_t)i * dim + c] += t * P[(size_t)k * dim + c];
            }
        for (int i = 0; i < dim; ++i)
            for (int c = 0; c < dim; ++c) {
                double a = 0.0;
                for (int k = 0; k < dim; ++k) a += TP[(size_t)i * dim + k] * T[(size_t)c * dim + k];
                P2[(size_t)i * dim + c] = a;
            }
        std::vector<double> RK((size_t)rows * dim, 0.0);   // R K^T
        for (int r = 0; r < rows; ++r)
            for (int q = 0; q < rows; ++q) {
                const double rr = R[(size_t)r * rows + q];
                if (rr == 0.0) continue;
                for (int c = 0; c < dim; ++c) RK[(size_t)r * dim + c] += rr * K[(size_t)c * rows + q];
            }
        for (int i = 0; i < dim; ++i)
            for (int r = 0; r < rows; ++r) {
                const double k = K[(size_t)i * rows + r];
                if (k == 0.0) continue;
                for (int c = 0; c < dim; ++c) P2[(size_t)i * dim + c] += k * RK[(size_t)r * dim + c];
            }
        P.swap(P2);
        return LIMU_OK;
    }
};

extern "C" {

void limu_ekf_default_params(limu_ekf_params *p) {   // odom_run.cpp:19-34 defaults where it sets them
    if (!p) return;
    memset(p, 0, sizeof *p);
    p->lidar_pose_trail = 20;
    p->noise_scale = 1.0;
    p->init_pos_noise = p->init_vel_noise = p->init_ori_noise = p->init_bga_noise = p->init_baa_noise = p->init_bat_noise = 1e-3;
    p->acc_process_noise = 0.03; p->gyro_process_noise = 0.00017;
    p->acc_process_noise_rev = 0.03; p->gyro_process_noise_rev = 0.00017;
    p->init_lidar_imu_time_noise = 1e-3; p->init_pos_trail_noise = 1e-3; p->init_ori_trail_noise = 1e-3; p->visualZuptR = 1e-3;
}

int limu_ekf_create(const limu_ekf_params *prm, limu_ekf **out) {   // EKF::EKF, ekf.cpp:63-190
    LIMU_REQUIRE(prm && out && prm->lidar_pose_trail >= 1 && prm->lidar_pose_trail <= 256, "limu_ekf_create: bad arguments (1 <= lidar_pose_trail <= 256)");
    limu_ekf *e = new limu_ekf;
    e->prm = *prm;
    e->trail = prm->lidar_pose_trail;
    e->dim = INNER + e->trail * POSE_DIM;
    e->noise_scale = prm->noise_scale * prm->noise_scale;
    const int n = e->dim;
    e->m.assign(n, 0.0);
    e->P.assign((size_t)n * n, 0.0);
    e->Q.assign(Q_DIM * Q_DIM, 0.0);
    e->dydx.assign(INNER * INNER, 0.0);
    e->dydq.assign(INNER * Q_DIM, 0.0);
    e->m[ORI] = 1.0; e->m[ROT_IMU_LIDAR] = 1.0;
    e->m[BAT] = e->m[BAT + 1] = e->m[BAT + 2] = 1.0;
    // initialize_process_covariance, ekf.cpp:580-617
    auto diag = [&](int at, int cnt, double v) { for (int i = 0; i < cnt; ++i) e->Pm(at + i, at + i) = v; };
    diag(POS, 3, sq(prm->init_pos_noise)); diag(VEL, 3, sq(prm->init_vel_noise)); diag(ORI, 4, 1.0);
    diag(BGA, 3, sq(prm->init_bga_noise)); diag(BAA, 3, sq(prm->init_baa_noise)); diag(BAT, 3, sq(prm->init_bat_noise));
    diag(GRAV, 3, sq(prm->init_lidar_imu_time_noise)); diag(POS_IMU_LIDAR, 3, sq(prm->init_pos_noise)); diag(ROT_IMU_LIDAR, 4, 1.0);
    e->Pm(SFT, SFT) = sq(prm->init_lidar_imu_time_noise);
    for (int k = 0; k < e->trail; ++k) { diag(LIDAR + k * POSE_DIM, 3, sq(prm->init_pos_trail_noise)); diag(LIDAR + k * POSE_DIM + 3, 4, sq(prm->init_ori_trail_noise)); }
    for (double &v : e->P) v *= e->noise_scale;
    for (int i = 0; i < 3; ++i) {
        e->Q[(Q_ACC + i) * Q_DIM + Q_ACC + i] = sq(prm->acc_process_noise) * e->noise_scale;
        e->Q[(Q_GYRO + i) * Q_DIM + Q_GYRO + i] = sq(prm->gyro_process_noise) * e->noise_scale;
    }
    *out = e;
    return LIMU_OK;
}

void limu_ekf_destroy(limu_ekf *e) { delete e; }

int limu_ekf_state_dim(limu_ekf *e, int32_t *dim) {
    LIMU_REQUIRE(e && dim, "limu_ekf_state_dim: null argument");
    *dim = e->dim;
    return LIMU_OK;
}

int limu_ekf_get_state(limu_ekf *e, double *m, double *P, double *time_out) {
    LIMU_REQUIRE(e, "limu_ekf_get_state: null handle");
    if (m) memcpy(m, e->m.data(), sizeof(double) * e->dim);
    if (P) memcpy(P, e->P.data(), sizeof(double) * (size_t)e->dim * e->dim);
    if (time_out) *time_out = e->first_sampleT + e->time;   // get_current_time, ekf.cpp:766-769
    return LIMU_OK;
}

int limu_ekf_set_state(limu_ekf *e, const double *m, const double *P) {
    LIMU_REQUIRE(e, "limu_ekf_set_state: null handle");
    if (m) memcpy(e->m.data(), m, sizeof(double) * e->dim);
    if (P) memcpy(e->P.data(), P, sizeof(double) * (size_t)e->dim * e->dim);
    return LIMU_OK;
}

// initialize_imu_global_orientation, ekf.cpp:194-211 AS EVIDENTLY INTENDED: the original writes four coefficients into a Vector3d and assigns
// it to a 4-segment (undefined behaviour, so there is nothing to be bit-faithful to): ORI = (w, x, y, z) of FromTwoVectors(calc_grav, xa),
// GRAV = calc_grav, P(ORI block) = diag(1, 1, 1, 0) * init_ori_noise^2 * noise_scale.
int limu_ekf_initialize_orientation(limu_ekf *e, const double xa[3], const double calc_grav[3]) {
    LIMU_REQUIRE(e && xa && calc_grav, "limu_ekf_initialize_orientation: null argument");
    // Eigen::Quaternion::setFromTwoVectors (Quaternion.h): v0 = a.normalized(), v1 = b.normalized(), c = v0.v1
    double a[3], b[3];
    const double na = sqrt((calc_grav[0] * calc_grav[0] + calc_grav[1] * calc_grav[1]) + calc_grav[2] * calc_grav[2]);
    const double nb = sqrt((xa[0] * xa[0] + xa[1] * xa[1]) + xa[2] * xa[2]);
    LIMU_REQUIRE(na > 0 && nb > 0, "limu_ekf_initialize_orientation: zero vector");
    for (int i = 0; i < 3; ++i) { a[i] = calc_grav[i] / na; b[i] = xa[i] / nb; }
    const double c = (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2];
    double q[4];   // w x y z
    if (c < -1.0 + 1e-12) {   // opposite vectors: any axis orthogonal to a (Eigen takes it from an SVD; the choice is not unique)
        double ax[3] = {1, 0, 0};
        if (fabs(a[0]) > 0.9) { ax[0] = 0; ax[1] = 1; }
        double o[3] = {a[1] * ax[2] - a[2] * ax[1], a[2] * ax[0] - a[0] * ax[2], a[0] * ax[1] - a[1] * ax[0]};
        const double no = sqrt((o[0] * o[0] + o[1] * o[1]) + o[2] * o[2]);
        q[0] = 0; q[1] = o[0] / no; q[2] = o[1] / no; q[3] = o[2] / no;
    } else {
        const double ax[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
        const double s = sqrt((1.0 + c) * 2.0), invs = 1.0 / s;
        q[0] = s * 0.5; q[1] = ax[0] * invs; q[2] = ax[1] * invs; q[3] = ax[2] * invs;
    }
    for (int i = 0; i < 4; ++i) e->m[ORI + i] = q[i];
    for (int i = 0; i < 3; ++i) e->m[GRAV + i] = calc_grav[i];
    for (int r = 0; r < 4; ++r)
        for (int c2 = 0; c2 < 4; ++c2) e->Pm(ORI + r, ORI + c2) = (r == c2 && r < 3) ? sq(e->prm.init_ori_noise) * e->noise_scale : 0.0;
    return LIMU_OK;
}

// EKF::predict, ekf.cpp:214-290 (+ calculate_S :471-484, extract_rot_dr helper.hpp:19-33, propagate_state :486-519, initialize_state_jacobians :521-578)
int limu_ekf_predict(limu_ekf *e, double t, const double xg[3], const double xa[3], const double calc_grav[3], const double trans_lidar_imu[3],
                     const double rot_lidar_imu[9]) {
    LIMU_REQUIRE(e && xg && xa && calc_grav && trans_lidar_imu && rot_lidar_imu, "limu_ekf_predict: null argument");
    double dt = 0.0;
    if (!e->first_sample) { dt = t - e->prev_sampleT; e->time = t - e->first_sampleT; }
    else { e->first_sampleT = t; e->first_sample = false; }
    e->prev_sampleT = t;
    if (dt <= 0.0) {
        if (e->time > 0) printf("Skipping KF predict, dt %g <=0.0\n", dt);
        return LIMU_OK;
    }
    const limu_ekf_params &p = e->prm;
    auto drift = [&](int at, double noise, double theta) {   // random-walk bias blocks of Q (:243-263)
        if (!(noise > 0.0)) return;
        double v = e->noise_scale * sq(noise);
        if (theta > 0.0) v *= (1 - exp(-2 * dt * theta)) / (2 * theta);
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) e->Q[(at + r) * Q_DIM + at + c] = r == c ? v : 0.0;
    };
    drift(Q_BGA_DRIFT, p.gyro_process_noise, p.gyro_process_noise_rev);
    drift(Q_BAA_DRIFT, p.acc_process_noise, p.acc_process_noise_rev);
    double *m = e->m.data();
    double A[16];
    calc_A(xg, m + BGA, dt, A);
    double qn[4];   // A * ORI
    for (int r = 0; r < 4; ++r) qn[r] = (A[4 * r] * m[ORI] + A[4 * r + 1] * m[ORI + 1]) + (A[4 * r + 2] * m[ORI + 2] + A[4 * r + 3] * m[ORI + 3]);
    double R[9], dR[4][9];
    quat_to_rot(qn, R);
    for (int i = 0; i < 4; ++i) {
        double ei[4] = {0, 0, 0, 0}, Ri[9];
        ei[i] = 1.0;
        quat_to_rot(ei, Ri);
        for (int k = 0; k < 9; ++k) dR[i][k] = Ri[k] - R[k];
    }
    // propagate_state
    for (int a = 0; a < 3; ++a) m[POS + a] += m[VEL + a] * dt;
    double T_ab[3];
    for (int a = 0; a < 3; ++a) T_ab[a] = m[BAT + a] * xa[a] - m[BAA + a];
    for (int a = 0; a < 3; ++a) m[VEL + a] += (((R[a] * T_ab[0] + R[3 + a] * T_ab[1]) + R[6 + a] * T_ab[2]) + m[GRAV + a]) * dt;   // R^T T_ab
    double prev_quat[4];
    for (int a = 0; a < 4; ++a) { prev_quat[a] = m[ORI + a]; m[ORI + a] = qn[a]; }
    if (p.acc_process_noise_rev > 0.0) { const double f = exp(-dt * p.acc_process_noise_rev); for (int a = 0; a < 3; ++a) m[BAA + a] *= f; }
    if (p.gyro_process_noise > 0.0) { const double f = exp(-dt * p.gyro_process_noise); for (int a = 0; a < 3; ++a) m[BGA + a] *= f; }
    for (int a = 0; a < 3; ++a) { m[GRAV + a] = calc_grav[a]; m[POS_IMU_LIDAR + a] = trans_lidar_imu[a]; }
    rot_to_quat(rot_lidar_imu, m + ROT_IMU_LIDAR);   // Quaterniond(rot).coeffs() = x y z w
    // initialize_state_jacobians
    double *Fx = e->dydx.data(), *Fw = e->dydq.data();
    auto FX = [&](int r, int c) -> double & { return Fx[r * INNER + c]; };
    auto FW = [&](int r, int c) -> double & { return Fw[r * Q_DIM + c]; };
    auto ident = [&](int at, int cnt, double v) { for (int r = 0; r < cnt; ++r) for (int c = 0; c < cnt; ++c) FX(at + r, at + c) = r == c ? v : 0.0; };
    ident(POS, 3, 1.0); ident(VEL, 3, 1.0);
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) FX(POS + r, VEL + c) = r == c ? dt : 0.0;
    ident(BGA, 3, 1.0); ident(BAA, 3, 1.0); ident(BAT, 3, 1.0); ident(GRAV, 3, 1.0); ident(POS_IMU_LIDAR, 3, 1.0); ident(ROT_IMU_LIDAR, 4, 1.0);
    double VO[12];   // Fx(VEL, ORI) before the product with A: column i = dR[i]^T T_ab dt
    for (int i = 0; i < 4; ++i)
        for (int r = 0; r < 3; ++r) VO[r * 4 + i] = ((dR[i][r] * T_ab[0] + dR[i][3 + r] * T_ab[1]) + dR[i][6 + r] * T_ab[2]) * dt;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 4; ++c) {
            double v = 0.0;
            for (int k = 0; k < 4; ++k) v += VO[r * 4 + k] * A[4 * k + c];
            FX(VEL + r, ORI + c) = v;
        }
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) FX(ORI + r, ORI + c) = A[4 * r + c];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) FW(VEL + r, Q_ACC + c) = R[3 * c + r] * dt;   // R^T dt
    const double h = dt / 2;
    const double dS[3][16] = {{0, h, 0, 0, -h, 0, 0, 0, 0, 0, 0, h, 0, 0, -h, 0},
                              {0, 0, h, 0, 0, 0, 0, -h, -h, 0, 0, 0, 0, h, 0, 0},
                              {0, 0, 0, h, 0, 0, h, 0, 0, -h, 0, 0, -h, 0, 0, 0}};
    for (int k = 0; k < 3; ++k) {
        double sq4[4];
        for (int r = 0; r < 4; ++r) { double v = 0.0; for (int c = 0; c < 4; ++c) v += dS[k][4 * r + c] * prev_quat[c]; sq4[r] = v; }
        for (int r = 0; r < 4; ++r) { double v = 0.0; for (int c = 0; c < 4; ++c) v += A[4 * r + c] * sq4[c]; FW(ORI + r, Q_GYRO + k) = v; }
    }
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) { FW(BGA + r, Q_BGA_DRIFT + c) = r == c ? 1.0 : 0.0; FW(BAA + r, Q_BAA_DRIFT + c) = r == c ? 1.0 : 0.0; }
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            double v = 0.0;
            for (int k = 0; k < 4; ++k) v += FX(VEL + r, ORI + k) * FW(ORI + k, Q_GYRO + c);
            FW(VEL + r, Q_GYRO + c) = v;
        }
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) FX(VEL + r, BGA + c) = -FW(VEL + r, Q_GYRO + c);
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 3; ++c) FX(ORI + r, BGA + c) = -FW(ORI + r, Q_GYRO + c);
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) { FX(VEL + r, BAA + c) = -R[3 * c + r] * dt; FX(VEL + r, BAT + c) = R[3 * c + r] * xa[c] * dt; }
    // covariance (:283-289): inner block, then the two off-diagonal strips against the pose trail
    const int n = e->dim, tail = n - INNER;
    std::vector<double> FP(INNER * INNER, 0.0), Pin(INNER * INNER, 0.0), FQ(INNER * Q_DIM, 0.0);
    for (int r = 0; r < INNER; ++r)
        for (int k = 0; k < INNER; ++k) {
            const double f = Fx[r * INNER + k];
            if (f == 0.0) continue;
            for (int c = 0; c < INNER; ++c) FP[r * INNER + c] += f * e->P[(size_t)k * n + c];
        }
    for (int r = 0; r < INNER; ++r)
        for (int k = 0; k < Q_DIM; ++k) {
            const double f = Fw[r * Q_DIM + k];
            if (f == 0.0) continue;
            for (int c = 0; c < Q_DIM; ++c) FQ[r * Q_DIM + c] += f * e->Q[k * Q_DIM + c];
        }
    for (int r = 0; r < INNER; ++r)
        for (int c = 0; c < INNER; ++c) {
            double v = 0.0, w = 0.0;
            for (int k = 0; k < INNER; ++k) v += FP[r * INNER + k] * Fx[c * INNER + k];
            for (int k = 0; k < Q_DIM; ++k) w += FQ[r * Q_DIM + k] * Fw[c * Q_DIM + k];
            Pin[r * INNER + c] = v + w;
        }
    std::vector<double> BL((size_t)tail * INNER, 0.0), TR((size_t)INNER * tail, 0.0);
    for (int r = 0; r < tail; ++r)
        for (int c = 0; c < INNER; ++c) {
            double v = 0.0;
            for (int k = 0; k < INNER; ++k) v += e->P[(size_t)(INNER + r) * n + k] * Fx[c * INNER + k];
            BL[(size_t)r * INNER + c] = v;
        }
    for (int r = 0; r < INNER; ++r)
        for (int k = 0; k < INNER; ++k) {
            const double f = Fx[r * INNER + k];
            if (f == 0.0) continue;
            for (int c = 0; c < tail; ++c) TR[(size_t)r * tail + c] += f * e->P[(size_t)k * n + INNER + c];
        }
    for (int r = 0; r < INNER; ++r) for (int c = 0; c < INNER; ++c) e->P[(size_t)r * n + c] = Pin[r * INNER + c];
    for (int r = 0; r < tail; ++r) for (int c = 0; c < INNER; ++c) e->P[(size_t)(INNER + r) * n + c] = BL[(size_t)r * INNER + c];
    for (int r = 0; r < INNER; ++r) for (int c = 0; c < tail; ++c) e->P[(size_t)r * n + INNER + c] = TR[(size_t)r * tail + c];
    return LIMU_OK;
}

int limu_ekf_normalize_quaternions(limu_ekf *e, int only_current) {
    LIMU_REQUIRE(e, "limu_ekf_normalize_quaternions: null handle");
    e->normalize_quaternions(only_current != 0);
    return LIMU_OK;
}

// zero_vel_update, ekf.cpp:657-678: H = [0 I] over the first VEL + 3 state entries, y = 0, R = r * noise_scale * I
int limu_ekf_zero_velocity_update(limu_ekf *e, double r) {
    LIMU_REQUIRE(e, "limu_ekf_zero_velocity_update: null handle");
    if (e->time - e->ZUPTtime < 0.25) return LIMU_OK;
    e->ZUPTtime = e->time;
    e->was_stationary = true;
    const int l = VEL + 3;
    std::vector<double> H(3 * l, 0.0), y(3, 0.0), R(9, 0.0);
    for (int i = 0; i < 3; ++i) { H[i * l + VEL + i] = 1.0; R[i * 3 + i] = r * e->noise_scale; }
    return e->update(3, l, H, y, R, false);
}

// update_visual_pose_aug, ekf.cpp:700-734: shift the pose trail by one (dropping the oldest), clone the current pose into slot 0 through a
// measurement update that ties slot 0 to POS / ORI (R = 1e-9 * noise_scale), Joseph form, symmetrise, normalise.
int limu_ekf_augment_pose_trail(limu_ekf *e) {
    LIMU_REQUIRE(e, "limu_ekf_augment_pose_trail: null handle");
    const int n = e->dim;
    std::vector<int> src(n);
    for (int i = 0; i < n; ++i) src[i] = i < LIDAR ? i : (i < LIDAR + POSE_DIM ? -1 : i - POSE_DIM);
    e->remap(src);
    for (int i = 0; i < 3; ++i) e->Pm(LIDAR + i, LIDAR + i) += sq(e->prm.init_pos_trail_noise) * e->noise_scale;
    for (int i = 3; i < POSE_DIM; ++i) e->Pm(LIDAR + i, LIDAR + i) += sq(e->prm.init_ori_trail_noise) * e->noise_scale;
    std::vector<double> H((size_t)POSE_DIM * n, 0.0), y(POSE_DIM, 0.0), R(POSE_DIM * POSE_DIM, 0.0);
    for (int i = 0; i < 3; ++i) { H[(size_t)i * n + POS + i] = 1.0; H[(size_t)i * n + LIDAR + i] = -1.0; }
    for (int i = 0; i < 4; ++i) { H[(size_t)(3 + i) * n + ORI + i] = 1.0; H[(size_t)(3 + i) * n + LIDAR + 3 + i] = -1.0; }
    for (int i = 0; i < POSE_DIM; ++i) R[i * POSE_DIM + i] = 1e-9 * e->noise_scale;
    LIMU_TRY(e->update(POSE_DIM, n, H, y, R, true));
    e->symmetrize();
    e->normalize_quaternions(false);
    e->augment_times.push_back(e->first_sampleT + e->time);   // (:727-731, as written: the count only grows once it exceeds the trail length)
    if (e->augment_count > e->trail) e->augment_count++;
    else e->augment_times.erase(e->augment_times.begin());
    return LIMU_OK;
}

// update_undo_augmentation, ekf.cpp:736-756: drop the newest pose of the trail (stationary device: the trail must not collapse into copies)
int limu_ekf_undo_augmentation(limu_ekf *e) {
    LIMU_REQUIRE(e, "limu_ekf_undo_augmentation: null handle");
    const int n = e->dim;
    std::vector<int> src(n);
    for (int i = 0; i < n; ++i) src[i] = i < LIDAR ? i : (i + POSE_DIM < n ? i + POSE_DIM : -1);
    e->remap(src);
    if (!e->augment_times.empty()) e->augment_times.pop_back();   // (the reference pops unconditionally: undefined on its empty vector)
    e->augment_count--;
    e->symmetrize();
    e->normalize_quaternions(false);
    return LIMU_OK;
}

// update_and_propagate, ekf.cpp:680-698
int limu_ekf_update_and_propagate(limu_ekf *e) {
    LIMU_REQUIRE(e, "limu_ekf_update_and_propagate: null handle");
    const double *v = &e->m[VEL];
    const double speed = sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]);
    if (fabs(speed) < 1e-3) {
        LIMU_TRY(limu_ekf_zero_velocity_update(e, e->prm.visualZuptR));
        LIMU_TRY(limu_ekf_undo_augmentation(e));
    }
    return limu_ekf_augment_pose_trail(e);
}

// The coupling the reference's design document promised and its code never wired: the registration result as a measurement of
// the filter. pose = lidar::KissICP's new pose {qx,qy,qz,qw, tx,ty,tz} taken as a direct observation of POS and ORI (w x y z, sign aligned
// with the state's quaternion), R = diag(pos_sigma^2 x3, ori_sigma^2 x4) * noise_scale, the same update() as the zero-velocity update.
// No counterpart in the reference: parity unpinned; tests check it against the textbook Kalman update in numpy.
int limu_ekf_update_lidar_pose(limu_ekf *e, const double pose[7], double pos_sigma, double ori_sigma) {
    LIMU_REQUIRE(e && pose && pos_sigma > 0 && ori_sigma > 0, "limu_ekf_update_lidar_pose: bad arguments");
    const int l = ORI + 4;
    std::vector<double> H((size_t)7 * l, 0.0), y(7), R(49, 0.0);
    double q[4] = {pose[3], pose[0], pose[1], pose[2]};   // w x y z
    const double *s = &e->m[ORI];
    if ((q[0] * s[0] + q[1] * s[1]) + (q[2] * s[2] + q[3] * s[3]) < 0) for (double &c : q) c = -c;   // q and -q are the same rotation
    for (int i = 0; i < 3; ++i) { H[(size_t)i * l + POS + i] = 1.0; y[i] = pose[4 + i]; R[i * 7 + i] = sq(pos_sigma) * e->noise_scale; }
    for (int i = 0; i < 4; ++i) { H[(size_t)(3 + i) * l + ORI + i] = 1.0; y[3 + i] = q[i]; R[(3 + i) * 7 + 3 + i] = sq(ori_sigma) * e->noise_scale; }
    return e->update(7, l, H, y, R, false);
}

}  // extern "C"
