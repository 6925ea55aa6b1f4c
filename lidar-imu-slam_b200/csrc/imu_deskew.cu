// imu_deskew.cu -- IMU-propagated backward deskew: the per-point loop of kalman::EKF::motion_compensation_with_imu
// (L/src/kalman/ekf.cpp:420-468). The host IMU forward pass (:315-410) stays where it is (it walks
// ~20 IMU samples through the EKF state) and hands over its pose table (kalman::Pose6D rows, ekf.hpp:88-105), the
// scan-end rotation / lidar position (:393-418) and the IMU->lidar lever arm; every point is then independent:
//   head   = the latest table row (excluding the last) whose offset_time is below the point's time (:422-433)
//   R_i    = R_head * AngleAxis(dt |w|, w/|w|)                                   (:444, helper.hpp:35-40)
//   T_ei   = p + v dt + 0.5 a dt^2 + R_i p_IL - p_lidar_end                      (:445)
//   P_comp = R_end^T (R_i P_i + T_ei), stored back as FLOAT and widened          (:448-468)
// The reference's serial loop never consumes the first point of the scan (:455-456), so that point is compensated again
// by every older head below its time; the kernel reproduces that on thread 0.
#include <algorithm>

#include "common.cuh"

namespace limu {

constexpr int IMU_MAX_ROWS = 96;

struct ImuDeskewArgs {
    unsigned char *rec;      // point records, `stride` bytes apart: float x,y,z at offset 0, float curvature (ms) at `toff`
    int stride, toff;
    int64_t n;
    const double *table;     // M x 22
    int M;
    double rot_end[9], pos_lidar_end[3], p_il[3];
    double *out;             // n x 3 (optional)
};

// Eigen::AngleAxisd(dt*|w|, w.normalized()).toRotationMatrix() (Eigen/src/Geometry/AngleAxis.h)
__device__ __forceinline__ void ang_vel_to_rmat(const double *w, double dt, double *R) {
    const double z = sqnorm3(w[0], w[1], w[2]);
    const double nrm = sqrt(z);
    double ax = w[0], ay = w[1], az = w[2];
    if (z > 0.0) { ax = w[0] / nrm; ay = w[1] / nrm; az = w[2] / nrm; }
    const double angle = dt * nrm;
    const double s = sin(angle), c = cos(angle);
    const double sx = s * ax, sy = s * ay, sz = s * az;
    const double cx = (1.0 - c) * ax, cy = (1.0 - c) * ay, cz = (1.0 - c) * az;
    double tmp;
    tmp = cx * ay; R[1] = tmp - sz; R[3] = tmp + sz;
    tmp = cx * az; R[2] = tmp + sy; R[6] = tmp - sy;
    tmp = cy * az; R[5] = tmp - sx; R[7] = tmp + sx;
    R[0] = cx * ax + c; R[4] = cy * ay + c; R[8] = cz * az + c;
}

__device__ __forceinline__ void compensate(const ImuDeskewArgs &A, const double *T /* table row in shared memory */, double t, float *xyz) {
    const double dt = t - T[0];
    double Rw[9], Ri[9], t1[3], t2[3];
    ang_vel_to_rmat(T + 4, dt, Rw);
    mat3mul(T + 13, Rw, Ri);
    mat3vec(Ri, A.p_il, t1);
    const double P[3] = {(double)xyz[0], (double)xyz[1], (double)xyz[2]};
    mat3vec(Ri, P, t2);
    double v[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const double Tei = (((T[10 + a] + T[7 + a] * dt) + (0.5 * T[1 + a]) * (dt * dt)) + t1[a]) - A.pos_lidar_end[a];
        v[a] = t2[a] + Tei;
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) xyz[a] = (float)((A.rot_end[a] * v[0] + A.rot_end[3 + a] * v[1]) + A.rot_end[6 + a] * v[2]);
}

static __global__ void __launch_bounds__(256) k_deskew_imu(const ImuDeskewArgs A) {
    __shared__ double tab[IMU_MAX_ROWS * 22];
    for (int k = threadIdx.x; k < A.M * 22; k += blockDim.x) tab[k] = A.table[k];
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n) return;
    float *q = reinterpret_cast<float *>(A.rec + (size_t)i * A.stride);
    float xyz[3] = {q[0], q[1], q[2]};
    const double t = (double)*reinterpret_cast<const float *>(A.rec + (size_t)i * A.stride + A.toff) / 1000.0;   // curvature / double(1000)
    int h = A.M - 2;
    while (h >= 0 && !(t > tab[22 * h])) --h;
    if (h >= 0) {
        compensate(A, tab + 22 * h, t, xyz);
        if (i == 0)   // the serial loop re-visits the first point under every older head (:455-456)
            for (int g = h - 1; g >= 0; --g) if (t > tab[22 * g]) compensate(A, tab + 22 * g, t, xyz);
        q[0] = xyz[0]; q[1] = xyz[1]; q[2] = xyz[2];
    }
    if (A.out) { A.out[3 * i] = (double)xyz[0]; A.out[3 * i + 1] = (double)xyz[1]; A.out[3 * i + 2] = (double)xyz[2]; }
}

}  // namespace limu

using namespace limu;

extern "C" int limu_deskew_imu(limu_ctx *c, void *points, int32_t stride_bytes, int32_t curvature_offset_bytes, int64_t n, const limu_imu_pose *table,
                               int32_t n_poses, const double rot_end[9], const double pos_lidar_end[3], const double p_imu_lidar[3], double *out_xyz,
                               int32_t write_back) {
    LIMU_TRY(bind(c));
    LIMU_REQUIRE(points && table && rot_end && pos_lidar_end && p_imu_lidar && n >= 1 && n_poses >= 2 && n_poses <= IMU_MAX_ROWS && stride_bytes >= 16 &&
                     stride_bytes % 4 == 0 && curvature_offset_bytes >= 12 && curvature_offset_bytes + 4 <= stride_bytes && curvature_offset_bytes % 4 == 0,
                 "limu_deskew_imu: bad arguments (need n >= 1, 2 <= n_poses <= 96, float x,y,z at offset 0 and a float time field inside each record)");
    static_assert(sizeof(limu_imu_pose) == 22 * sizeof(double), "limu_imu_pose layout");
    LIMU_TRY(stage_in(c, c->in1, points, (size_t)n * stride_bytes));
    LIMU_TRY(stage_in(c, c->tmp0, table, (size_t)n_poses * sizeof(limu_imu_pose)));
    if (out_xyz) LIMU_TRY(c->out0.reserve((size_t)n * 24, c->stream));
    ImuDeskewArgs A;
    A.rec = c->in1.as<unsigned char>(); A.stride = stride_bytes; A.toff = curvature_offset_bytes; A.n = n;
    A.table = c->tmp0.as<double>(); A.M = n_poses;
    for (int k = 0; k < 9; ++k) A.rot_end[k] = rot_end[k];
    for (int k = 0; k < 3; ++k) { A.pos_lidar_end[k] = pos_lidar_end[k]; A.p_il[k] = p_imu_lidar[k]; }
    A.out = out_xyz ? c->out0.as<double>() : nullptr;
    k_deskew_imu<<<div_up(n, 256), 256, 0, c->stream>>>(A);
    LIMU_LAUNCHED();
    if (out_xyz) LIMU_CUDA_TRY(cudaMemcpyAsync(out_xyz, c->out0.p, (size_t)n * 24, cudaMemcpyDeviceToHost, c->stream));
    if (write_back)   // only the 12 bytes of x,y,z of every record travel back
        LIMU_CUDA_TRY(cudaMemcpy2DAsync(points, (size_t)stride_bytes, c->in1.p, (size_t)stride_bytes, 12, (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    return check_status(c);
}
