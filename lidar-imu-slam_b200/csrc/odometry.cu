// odometry.cu -- lidar::KissICP (L/include/limu/sensors/lidar/icp.hpp:31-68, L/src/sensors/lidar/icp.cpp)
// as a device pipeline: one H2D of the raw scan, every stage enqueued on the handle's stream with
// device-side counts, ONE host synchronisation per scan (to read the pose the scalar host glue needs:
// adaptive threshold, constant-velocity prediction -- L/src/sensors/lidar/helpers/threshold.cpp, icp.cpp:138-163).
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "frame_fusion.cuh"
#include "ops.cuh"
#include "voxel_map.cuh"

namespace limu {
int icp_device(limu_map *m, const double *points_dev, double *work_dev, int64_t n_max, const int *n_dev, const double *init_pose_host,
               double tau, double th, int max_iter, double eps, double *partials_dev, size_t partial_rows, double *out13_dev, int64_t n_hint,
               double *est_trace_dev, long long *ncorr_trace_dev, double *hg_trace_dev, int max_iter_all_ranks, const FrameFusion *fuse, int icp_mode);
int icp_partial_rows(limu_ctx *c);

// theta = Eigen::AngleAxisd(model_dev.rotationMatrix()).angle() (threshold.cpp:7): quaternion -> matrix
// (Eigen Quaternion::toRotationMatrix) -> quaternion (Eigen RotationBase assign, Quaternion.h) -> 2 atan2(|v|, |w|).
static double rotation_angle(const Pose &T) {
    const double x = T.qx, y = T.qy, z = T.qz, w = T.qw;
    const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
    double m[3][3];
    m[0][0] = 1 - (tyy + tzz); m[0][1] = txy - twz; m[0][2] = txz + twy;
    m[1][0] = txy + twz; m[1][1] = 1 - (txx + tzz); m[1][2] = tyz - twx;
    m[2][0] = txz - twy; m[2][1] = tyz + twx; m[2][2] = 1 - (txx + tyy);
    double c[4];
    double t = (m[0][0] + m[1][1]) + m[2][2];
    if (t > 0) {
        t = sqrt(t + 1.0);
        c[3] = 0.5 * t; t = 0.5 / t;
        c[0] = (m[2][1] - m[1][2]) * t; c[1] = (m[0][2] - m[2][0]) * t; c[2] = (m[1][0] - m[0][1]) * t;
    } else {
        int i = 0;
        if (m[1][1] > m[0][0]) i = 1;
        if (m[2][2] > m[i][i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = sqrt(m[i][i] - m[j][j] - m[k][k] + 1.0);
        c[i] = 0.5 * t; t = 0.5 / t;
        c[3] = (m[k][j] - m[j][k]) * t; c[j] = (m[j][i] + m[i][j]) * t; c[k] = (m[k][i] + m[i][k]) * t;
    }
    const double n = sqrt(sqnorm3(c[0], c[1], c[2]));
    return n != 0.0 ? 2.0 * atan2(n, fabs(c[3])) : 0.0;
}
}  // namespace limu

struct limu_odom {
    limu_ctx *ctx = nullptr;
    limu_odom_config cfg;
    limu_map *map = nullptr;
    std::vector<limu::Pose> poses;
    // AdaptiveThreshold (helpers/threshold.hpp:9-33)
    double model_error_sq = 0.0;
    int num_samples = 0;
    limu::Pose model_deviation = limu::pose_identity();
    // device buffers
    limu::DevBuf raw, ts, frame, down, src0, src, work, world, partials, d2, res;
    int vox_word = 0;                       // which of its two status words the last k_voxelize of this handle reports into
    limu::VoxelizeScratch vx;
    limu::PreScratch pre;                   // limu_odom_register_msg: frame::Lidar::process_frame on the device
    int64_t nk_hint = 2048, nd_hint = 16384;   // keypoints / downsampled points of the previous scan (launch shapes; the kernels take any count)
    // limu_odom_prefetch: the next scan is uploaded on its own stream while the current one is being registered
    limu::DevBuf pf_buf[2];                 // two slots: the scan about to be registered and the one after it
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t pf_done[2] = {nullptr, nullptr};
    const void *pf_host[2] = {nullptr, nullptr};
    int64_t pf_n[2] = {-1, -1};
    uint64_t pf_seq[2] = {0, 0}, pf_counter = 0;   // order in which the pending uploads were requested
    // Speculative voxelize (LIMU_OPT_SPECULATE, on by default): when the library knows which scan comes next (a device-pointer hint from
    // limu_odom_hint_next_dev, or the scan limu_odom_prefetch is uploading) that scan's deskew + downsampling launch is enqueued right
    // behind this scan's frame kernel, with its deskew twist left on the device by that kernel, so the host round trip of this scan
    // (result copy, wake-up, scalar glue, launch) overlaps it instead of idling the GPU.
    bool speculate = true;
    bool cluster_loop = false;              // LIMU_OPT_CLUSTER_LOOP: run the Gauss-Newton loop on one 16-CTA cluster (registration.cu, k_frame_cluster); measured slower, opt-in
    const void *hint_ptr = nullptr;         // next scan (device float4 rows), set by the caller before registering the current one
    int64_t hint_n = 0;
    const void *spec_ptr = nullptr;         // scan whose k_voxelize is already in flight / done
    int64_t spec_n = 0;
    int spec_deskewed = 0;
    int spec_slot = -1;                     // prefetch slot the speculative launch reads (-1: a caller-owned device buffer)
    limu::DevBuf twist_next;                // 6 doubles written by the frame kernel
    cudaEvent_t frame_done = nullptr;       // recorded after the result copy of the current scan
    cudaEvent_t spec_launched = nullptr;    // recorded right behind the speculative launch (completes when that kernel has finished)
    cudaEvent_t spec_ev[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // device time of the speculative launch (two pairs, alternating), folded into LIMU_STAGE_DOWNSAMPLE one call later
    bool spec_timed = false;
    int spec_par = 0;
    // This scan's clouds leave on the copy stream while the speculative launch runs, so `down` is double-buffered (the launch writes the
    // buffer this scan does not use); `src` is only rewritten by the NEXT frame kernel, which is launched after the clouds have arrived.
    limu::DevBuf down_alt;
    int down_cur = 0;
    cudaEvent_t clouds_done = nullptr;
};

using namespace limu;

static bool odom_has_moved(const limu_odom *o) {   // icp.cpp:156-163
    if (o->poses.empty()) return false;
    const Pose d = mul(inverse(o->poses.front()), o->poses.back());
    return sqrt(sqnorm3(d.tx, d.ty, d.tz)) > 5.0 * o->cfg.min_motion_th;
}
static double odom_compute_threshold(limu_odom *o) {   // threshold.cpp:16-28 (+ :5-12)
    const double theta = rotation_angle(o->model_deviation);
    const double delta_rot = 2.0 * o->cfg.max_range * sin(theta / 2.0);
    const double delta_trans = sqrt(sqnorm3(o->model_deviation.tx, o->model_deviation.ty, o->model_deviation.tz));
    const double model_error = delta_rot + delta_trans;
    if (model_error > o->cfg.min_motion_th) { o->model_error_sq += model_error * model_error; o->num_samples++; }
    if (o->num_samples < 1) return o->cfg.initial_threshold;
    return sqrt(o->model_error_sq / o->num_samples);
}
static double odom_adaptive_threshold(limu_odom *o) {   // icp.cpp:138-144
    if (!odom_has_moved(o)) return o->cfg.initial_threshold;
    return odom_compute_threshold(o);
}
static Pose odom_prediction(const limu_odom *o) {   // icp.cpp:146-154
    const size_t N = o->poses.size();
    if (N < 2) return pose_identity();
    return mul(inverse(o->poses[N - 2]), o->poses[N - 1]);
}

static int odom_side_stream(limu_odom *o) {
    if (!o->copy_stream) {
        LIMU_CUDA_TRY(cudaStreamCreateWithFlags(&o->copy_stream, cudaStreamNonBlocking));
        for (int sl = 0; sl < 2; ++sl) LIMU_CUDA_TRY(cudaEventCreateWithFlags(&o->pf_done[sl], cudaEventDisableTiming));
    }
    return LIMU_OK;
}

// Everything after the scan is in device memory.
// Input already in device memory. mode 0: float4 {x,y,z,t}; 1: records `stride` bytes apart + FP64 timestamps; 2: double xyz.
static int odom_register_device(limu_odom *o, const void *raw_dev, int mode, int stride, const double *ts_dev, int64_t n, double pose_out[7], double *down_xyz,
                                int64_t *n_down, double *keypoints_xyz, int64_t *n_keypoints, limu_frame_stats *stats) {
    limu_ctx *c = o->ctx;
    // deskew gate (icp.cpp:40-46): config.deskew && poses.size() > 2; twist = delta_pose(poses[N-2], poses[N-1]) (deskew.cpp:14)
    const size_t NP = o->poses.size();
    const int deskewed = (mode != 2 && o->cfg.deskew && NP > 2) ? 1 : 0;
    double twist[6] = {0, 0, 0, 0, 0, 0};
    if (deskewed) se3_log(mul(inverse(o->poses[NP - 2]), o->poses[NP - 1]), twist);
    // was this scan's k_voxelize already enqueued behind the previous scan? Its outputs live in frame / src0 / the other `down` buffer:
    // growing those buffers for a larger hinted scan must then keep their contents.
    const bool spec_hit = mode == 0 && n > 0 && o->spec_ptr && raw_dev == o->spec_ptr && n == o->spec_n && o->spec_deskewed == deskewed;
    if (o->spec_ptr && !spec_hit && o->spec_slot >= 0 && o->pf_buf[o->spec_slot].p == o->spec_ptr) {
        // the caller registered something else than the scan it had prefetched: that upload is stale, do not speculate on it again
        o->pf_host[o->spec_slot] = nullptr; o->pf_n[o->spec_slot] = -1;
    }
    o->spec_ptr = nullptr; o->spec_slot = -1;
    if (spec_hit) o->down_cur ^= 1;   // the speculative launch wrote the other `down` buffer
    // the scan after this one, if the caller told us where it is: a device-pointer hint, or the OLDEST upload limu_odom_prefetch has pending
    const void *next_ptr = nullptr;
    int64_t next_n = 0;
    int next_slot = -1;
    cudaEvent_t next_ready = nullptr;
    if (o->speculate && mode == 0) {
        if (o->hint_ptr) { next_ptr = o->hint_ptr; next_n = o->hint_n; }
        else {
            for (int sl = 0; sl < 2; ++sl)
                if (o->pf_host[sl] && o->pf_n[sl] > 0 && (next_slot < 0 || o->pf_seq[sl] < o->pf_seq[next_slot])) next_slot = sl;
            if (next_slot >= 0) { next_ptr = o->pf_buf[next_slot].p; next_n = o->pf_n[next_slot]; next_ready = o->pf_done[next_slot]; }
        }
    }
    o->hint_ptr = nullptr; o->hint_n = 0;   // a hint is good for one call only
    const size_t nb = (size_t)std::max<int64_t>(std::max<int64_t>(n, next_n), 1) * 24;
    const bool keep = spec_hit;
    if (o->speculate || o->down_cur) LIMU_TRY(o->down_alt.reserve(nb, c->stream, keep));
    limu::DevBuf &down_mine = o->down_cur ? o->down_alt : o->down, &down_other = o->down_cur ? o->down : o->down_alt;
    LIMU_TRY(o->frame.reserve(nb, c->stream, keep));
    LIMU_TRY(o->down.reserve(nb, c->stream, keep));
    LIMU_TRY(o->src0.reserve(nb, c->stream, keep));
    LIMU_TRY(o->src.reserve(nb, c->stream));
    LIMU_TRY(o->work.reserve(nb, c->stream));
    LIMU_TRY(o->world.reserve(nb, c->stream));
    const int rows = icp_partial_rows(c);
    {   // per-CTA rows of the Gauss-Newton loop: 32 (value, stamp) word pairs per row cover both residual variants; never anything that looks like a stamp in a fresh buffer
        const void *before = o->partials.p;
        LIMU_TRY(o->partials.reserve((size_t)2 * rows * 32 * 16 + 256, c->stream));
        if (o->partials.p != before) LIMU_CUDA_TRY(cudaMemsetAsync(o->partials.p, 0, o->partials.bytes, c->stream));
    }
    // Per-HANDLE result block in device memory (two handles of one context may interleave their scans, and a speculative launch of one
    // must not land in the other's counts): 16 ints -- [0]=n_down [1]=n_src0 [2]=n_keypoints, [4..7] / [8..11] = the two status words
    // k_voxelize alternates between (its own words, so that a speculative launch is never blamed on the scan before it), [12..15] = the
    // frame kernel's status word -- then pose + loop statistics (13 doubles). ONE copy per scan.
    if (!o->res.p) {
        LIMU_TRY(o->res.reserve(32 * sizeof(double)));
        LIMU_CUDA_TRY(cudaMemsetAsync(o->res.p, 0, o->res.bytes, c->stream));
    }
    int *cnt = o->res.as<int>();
    DevStatus *vox_status = reinterpret_cast<DevStatus *>(cnt + 4);
    DevStatus *frame_status = reinterpret_cast<DevStatus *>(cnt + 12);
    double *out13 = o->res.as<double>() + 8;
    const double v = o->cfg.voxel_size;

    // deskew_scan + voxelize's two downsampling stages (icp.cpp:36-47, :126-131): one cooperative launch
    if (!spec_hit) {
        LIMU_TRY(prof_begin(c, LIMU_STAGE_DOWNSAMPLE));
        LIMU_TRY(voxelize_device(c, o->vx, raw_dev, mode, stride, ts_dev, deskewed, twist, n, v, o->frame.as<double>(), down_mine.as<double>(), o->src0.as<double>(), cnt + 0,
                                 nullptr, vox_status, &o->vox_word));
        LIMU_TRY(prof_end(c, LIMU_STAGE_DOWNSAMPLE));
    }
    // host scalar glue (icp.cpp:66-71)
    const double sigma = odom_adaptive_threshold(o);
    const Pose pred = odom_prediction(o);
    const Pose last = o->poses.empty() ? pose_identity() : o->poses.back();
    const Pose init = mul(last, pred);
    double init7[7];
    pose_store(init, init7);

    // iqr_processing (icp.cpp:133) + ICP (icp.cpp:74-76) + local_map.update(down_sampled, new_pose) (icp.cpp:81;
    // voxel_hash_map.cpp:138-144): ONE persistent cooperative launch (registration.cu)
    LIMU_TRY(map_maybe_grow(o->map, n));
    LIMU_TRY(o->map->pslot.reserve((size_t)std::max<int64_t>(n, 1) * 4, c->stream));
    LIMU_TRY(o->d2.reserve((size_t)std::max<int64_t>(n, 1) * 8, c->stream));
    FrameFusion fuse;
    fuse.iqr_in = o->src0.as<double>(); fuse.iqr_n = cnt + 1; fuse.iqr_d2 = o->d2.as<double>(); fuse.iqr_out = o->src.as<double>(); fuse.iqr_count = cnt + 2;
    fuse.upd_down = down_mine.as<double>(); fuse.upd_n = cnt + 0; fuse.upd_world = o->world.as<double>(); fuse.upd_pslot = o->map->pslot.as<unsigned int>();
    fuse.upd_birth_base = o->map->birth_base;
    const bool speculate = next_ptr && next_n > 0;
    const int next_deskew = (o->cfg.deskew && NP + 1 > 2) ? 1 : 0;   // the gate of icp.cpp:40-46 as the next scan will see it
    fuse.twist_out = nullptr;
    fuse.allow_cluster = o->cluster_loop ? 1 : 0;
    fuse.status = frame_status;
    pose_store(last, fuse.last_pose);
    if (speculate && next_deskew) {
        LIMU_TRY(o->twist_next.reserve(6 * sizeof(double), c->stream));
        fuse.twist_out = o->twist_next.as<double>();
    }
    const int64_t upper_before = o->map->used_upper;
    LIMU_TRY(icp_device(o->map, o->src.as<double>(), o->work.as<double>(), n, cnt + 2, init7, 3.0 * sigma, sigma / 3.0, o->cfg.icp_max_iteration,
                        o->cfg.estimation_threshold, o->partials.as<double>(), (size_t)rows, out13, o->nk_hint, nullptr, nullptr, nullptr, -1, &fuse, o->cfg.icp_mode));
    o->map->birth_base += (uint64_t)n;
    o->map->used_upper = upper_before + std::max<int64_t>(n, 0);   // safe bound until the exact count arrives (at most one new voxel per point)

    // the one synchronisation of the scan
    double *h = static_cast<double *>(c->h_pinned) + 32;
    const int my_vox_word = o->vox_word;   // the status word of THIS scan's k_voxelize (a speculative launch below moves on to the other one)
    LIMU_CUDA_TRY(cudaMemcpyAsync(h, o->res.p, (8 + 13) * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    // device time of the speculative launch that prepared THIS scan (recorded one call ago into the event pair `spec_par`): read after this
    // scan's sync, when it is certainly complete; the launch made below uses the other pair
    const int acct = o->spec_timed ? o->spec_par : -1;
    o->spec_timed = false;
    if (speculate) {
        if (!o->frame_done) {
            LIMU_CUDA_TRY(cudaEventCreateWithFlags(&o->frame_done, cudaEventDisableTiming));
            LIMU_CUDA_TRY(cudaEventCreateWithFlags(&o->spec_launched, cudaEventDisableTiming));
            for (int k = 0; k < 4; ++k) LIMU_CUDA_TRY(cudaEventCreate(&o->spec_ev[k >> 1][k & 1]));
        }
        LIMU_CUDA_TRY(cudaEventRecord(o->frame_done, c->stream));
        // stream order: frame kernel (writes twist_next) -> result copy -> [upload of the next scan done] -> k_voxelize of the next scan
        // (reads twist_next; overwrites frame, src0, the counts and its status word -- dead for this scan -- and the OTHER `down` buffer)
        if (next_ready) LIMU_CUDA_TRY(cudaStreamWaitEvent(c->stream, next_ready, 0));
        const int par = o->spec_par ^ 1;
        if (c->profiling) LIMU_CUDA_TRY(cudaEventRecord(o->spec_ev[par][0], c->stream));
        LIMU_TRY(voxelize_device(c, o->vx, next_ptr, 0, 0, nullptr, next_deskew, nullptr, next_n, v, o->frame.as<double>(), down_other.as<double>(),
                                 o->src0.as<double>(), cnt + 0, next_deskew ? o->twist_next.as<double>() : nullptr, vox_status, &o->vox_word));
        if (c->profiling) { LIMU_CUDA_TRY(cudaEventRecord(o->spec_ev[par][1], c->stream)); o->spec_timed = true; o->spec_par = par; }
        LIMU_CUDA_TRY(cudaEventRecord(o->spec_launched, c->stream));
        o->spec_ptr = next_ptr; o->spec_n = next_n; o->spec_deskewed = next_deskew; o->spec_slot = next_slot;
        LIMU_CUDA_TRY(cudaEventSynchronize(o->frame_done));   // wakes up when the frame kernel and the copy are done; k_voxelize keeps running
    } else {
        LIMU_CUDA_TRY(cudaStreamSynchronize(c->stream));
    }
    if (acct >= 0 && c->profiling) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, o->spec_ev[acct][0], o->spec_ev[acct][1]) == cudaSuccess) c->stage_ms[LIMU_STAGE_DOWNSAMPLE] += (double)ms;
        else (void)cudaGetLastError();
    }
    LIMU_TRY(prof_collect(c));
    const int *hc = reinterpret_cast<const int *>(h);
    const int64_t nd = hc[0], nk = hc[2];
    const double *ho = h + 8;
    const Pose new_pose = pose_load(ho);
    // Commit the whole frame -- map bookkeeping, pose history, threshold state -- BEFORE looking at the device status: the frame kernel
    // has already inserted this scan into the map (points with out-of-range voxel indices were left out by both downsampling and the
    // insert), so an error return must not leave a map that holds a scan without a pose.
    o->map->used_upper = upper_before + nd;   // exact: at most one new voxel per inserted point
    o->nk_hint = std::max<int64_t>(nk, 256);
    o->nd_hint = std::max<int64_t>(nd, 256);
    o->model_deviation = mul(inverse(init), new_pose);   // icp.cpp:78-79
    o->poses.push_back(new_pose);                        // :82
    if (pose_out) pose_store(new_pose, pose_out);
    if (n_down) *n_down = nd;
    if (n_keypoints) *n_keypoints = nk;
    bool copied = false;
    cudaStream_t cloud_stream = c->stream;
    if (speculate && ((down_xyz && nd > 0) || (keypoints_xyz && nk > 0))) {
        // the main stream is busy with the next scan's k_voxelize: this scan's clouds leave on the copy stream (everything they read was
        // complete at frame_done, and nothing in flight writes it)
        LIMU_TRY(odom_side_stream(o));
        if (!o->clouds_done) LIMU_CUDA_TRY(cudaEventCreateWithFlags(&o->clouds_done, cudaEventDisableTiming));
        cloud_stream = o->copy_stream;
    }
    if (down_xyz && nd > 0) { LIMU_CUDA_TRY(cudaMemcpyAsync(down_xyz, down_mine.p, (size_t)nd * 24, cudaMemcpyDeviceToHost, cloud_stream)); copied = true; }
    if (keypoints_xyz && nk > 0) { LIMU_CUDA_TRY(cudaMemcpyAsync(keypoints_xyz, o->src.p, (size_t)nk * 24, cudaMemcpyDeviceToHost, cloud_stream)); copied = true; }
    if (copied) {
        if (cloud_stream != c->stream) {
            LIMU_CUDA_TRY(cudaEventRecord(o->clouds_done, cloud_stream));
            LIMU_CUDA_TRY(cudaEventSynchronize(o->clouds_done));
        } else {
            LIMU_CUDA_TRY(cudaStreamSynchronize(c->stream));
        }
    }
    if (stats) {
        stats->n_points = n; stats->n_down = nd; stats->n_keypoints = nk; stats->sigma = sigma; stats->deskewed = deskewed; stats->reserved0 = spec_hit ? 1 : 0;
        stats->icp.iterations = (int)ho[7]; stats->icp.converged = (int)ho[8]; stats->icp.last_ncorr = (int64_t)ho[9];
        stats->icp.mean_candidates = nk > 0 ? ho[10] / (double)nk : 0.0;
        stats->icp.miss_fraction = nk > 0 ? ho[11] / (double)nk : 0.0;
    }
    {   // device status of this scan: its own k_voxelize word and the frame kernel's word (insert)
        DevStatus st, sv;
        memcpy(&st, hc + 12, sizeof st);
        memcpy(&sv, hc + 4 + 4 * my_vox_word, sizeof sv);
        if (st.key_range | st.table_full | st.pad[0]) LIMU_CUDA_TRY(cudaMemsetAsync(frame_status, 0, sizeof(DevStatus), c->stream));   // (the next frame kernel is launched after this)
        st.key_range |= sv.key_range; st.table_full |= sv.table_full;
        DevStatus none = {0, 0, {0, 0}};
        (void)none;
        if (st.key_range) { set_error("voxel index outside the packed key range (|index| >= 2^20) or NaN coordinate: the frame was registered without those points"); return LIMU_ERR_KEY_RANGE; }
        if (st.table_full) { set_error("voxel hash table full"); return LIMU_ERR_MAP_FULL; }
    }
    return LIMU_OK;
}

extern "C" {

void limu_odom_default_config(limu_odom_config *cfg) {   // lidar/frame.hpp:64-80
    if (!cfg) return;
    memset(cfg, 0, sizeof *cfg);
    cfg->max_range = 100.0;
    cfg->voxel_size = cfg->max_range / 100.0;
    cfg->max_points_per_voxel = 10;
    cfg->deskew = 0;
    cfg->min_motion_th = 0.1;
    cfg->icp_max_iteration = 500;
    cfg->initial_threshold = 2.0;
    cfg->estimation_threshold = 0.0001;
}

int limu_odom_create(limu_ctx *c, const limu_odom_config *cfg, limu_odom **out) {
    LIMU_TRY(bind(c));
    LIMU_REQUIRE(cfg && out, "limu_odom_create: null argument");
    LIMU_REQUIRE(cfg->voxel_size > 0 && cfg->max_points_per_voxel >= 1, "limu_odom_create: voxel_size must be > 0 and max_points_per_voxel >= 1");
    LIMU_REQUIRE((cfg->icp_mode & ~(LIMU_ICP_NN27 | LIMU_ICP_PLANE)) == 0, "limu_odom_create: unknown icp_mode bits");
    limu_odom *o = new limu_odom;
    o->ctx = c;
    o->cfg = *cfg;
    if (const char *e = getenv("LIMU_SPECULATE")) o->speculate = atoi(e) != 0;        // default of LIMU_OPT_SPECULATE (on)
    if (const char *e = getenv("LIMU_CLUSTER_LOOP")) o->cluster_loop = atoi(e) != 0;   // default of LIMU_OPT_CLUSTER_LOOP (off)
    int64_t capv = cfg->map_capacity_voxels;
    if (capv <= 0) {   // a sensor sees a shell, not a ball: ~ (2 r / v)^2 * 8 voxels is generous for one neighbourhood
        const double side = 2.0 * cfg->max_range / cfg->voxel_size;
        capv = (int64_t)std::min(4.0e6, std::max(65536.0, side * side * 8.0));
    }
    int st = limu_map_create(c, cfg->voxel_size, cfg->max_range, cfg->max_points_per_voxel, capv, &o->map);   // icp.hpp:35-37
    if (st != LIMU_OK) { delete o; return st; }
    *out = o;
    return LIMU_OK;
}

void limu_odom_destroy(limu_odom *o) {
    if (!o) return;
    cudaSetDevice(o->ctx->device);
    cudaStreamSynchronize(o->ctx->stream);
    limu_map_destroy(o->map);
    if (o->copy_stream) { cudaStreamSynchronize(o->copy_stream); cudaStreamDestroy(o->copy_stream); cudaEventDestroy(o->pf_done[0]); cudaEventDestroy(o->pf_done[1]); }
    DevBuf *bufs[] = {&o->res, &o->pf_buf[0], &o->pf_buf[1], &o->d2, &o->raw, &o->ts, &o->frame, &o->down, &o->src0, &o->src, &o->work, &o->world, &o->partials};
    for (auto *b : bufs) b->release();
    o->vx.release();
    o->pre.release();
    o->twist_next.release();
    o->down_alt.release();
    if (o->frame_done) { cudaEventDestroy(o->frame_done); cudaEventDestroy(o->spec_launched); for (int k = 0; k < 4; ++k) cudaEventDestroy(o->spec_ev[k >> 1][k & 1]); }
    if (o->clouds_done) cudaEventDestroy(o->clouds_done);
    delete o;
}

int limu_odom_register_frame(limu_odom *o, const float *xyzt, int64_t n, double pose_out[7], double *down_xyz, int64_t *n_down,
                             double *keypoints_xyz, int64_t *n_keypoints, limu_frame_stats *stats) {
    LIMU_REQUIRE(o && n >= 0 && (n == 0 || xyzt), "limu_odom_register_frame: bad arguments");
    LIMU_TRY(bind(o->ctx));
    int hit = -1;
    for (int s = 0; s < 2; ++s) if (n > 0 && o->pf_host[s] == xyzt && o->pf_n[s] == n) hit = s;
    if (hit >= 0 && o->spec_ptr && o->spec_ptr == o->pf_buf[hit].p && o->spec_n == n) {
        // uploaded ahead of time AND already through k_voxelize (enqueued behind the previous scan): register it where it lies
        o->pf_host[hit] = nullptr; o->pf_n[hit] = -1;
        return odom_register_device(o, o->pf_buf[hit].p, 0, 0, nullptr, n, pose_out, down_xyz, n_down, keypoints_xyz, n_keypoints, stats);
    }
    if (hit >= 0) {   // uploaded ahead of time by limu_odom_prefetch
        LIMU_CUDA_TRY(cudaStreamWaitEvent(o->ctx->stream, o->pf_done[hit], 0));
        std::swap(o->raw, o->pf_buf[hit]);
        o->pf_host[hit] = nullptr; o->pf_n[hit] = -1;
    } else {
        LIMU_TRY(stage_in(o->ctx, o->raw, xyzt, (size_t)n * 16));
    }
    return odom_register_device(o, o->raw.p, 0, 0, nullptr, n, pose_out, down_xyz, n_down, keypoints_xyz, n_keypoints, stats);
}

int limu_odom_prefetch(limu_odom *o, const float *xyzt, int64_t n) {
    LIMU_REQUIRE(o && n >= 0 && (n == 0 || xyzt), "limu_odom_prefetch: bad arguments");
    LIMU_TRY(bind(o->ctx));
    if (n == 0) return LIMU_OK;
    LIMU_TRY(odom_side_stream(o));
    for (int s = 0; s < 2; ++s) if (o->pf_host[s] == xyzt && o->pf_n[s] == n) return LIMU_OK;   // already in flight
    // a free slot, else the OLDEST pending upload is given up (its scan was evidently never registered)
    int s = o->pf_host[0] == nullptr ? 0 : (o->pf_host[1] == nullptr ? 1 : (o->pf_seq[0] < o->pf_seq[1] ? 0 : 1));
    if (o->spec_ptr && o->pf_buf[s].p == o->spec_ptr) {
        // the speculative k_voxelize of the scan in this slot may still be reading it on the main stream: let it finish, and forget the
        // speculation (the slot is about to hold a different scan, possibly of the same size)
        LIMU_CUDA_TRY(cudaEventSynchronize(o->spec_launched));
        o->spec_ptr = nullptr; o->spec_slot = -1;
    }
    o->pf_host[s] = nullptr; o->pf_n[s] = -1;
    if (o->pf_buf[s].bytes < (size_t)n * 16) {   // growing may free a buffer the copy stream still writes: drain it first
        LIMU_CUDA_TRY(cudaStreamSynchronize(o->copy_stream));
        LIMU_TRY(o->pf_buf[s].reserve((size_t)n * 16, o->copy_stream));
    }
    LIMU_CUDA_TRY(cudaMemcpyAsync(o->pf_buf[s].p, xyzt, (size_t)n * 16, cudaMemcpyHostToDevice, o->copy_stream));
    LIMU_CUDA_TRY(cudaEventRecord(o->pf_done[s], o->copy_stream));
    o->pf_host[s] = xyzt; o->pf_n[s] = n; o->pf_seq[s] = ++o->pf_counter;
    return LIMU_OK;
}

int limu_odom_register_cloud(limu_odom *o, const void *points, int32_t stride_bytes, const double *timestamps, int64_t n, double pose_out[7],
                             double *down_xyz, int64_t *n_down, double *keypoints_xyz, int64_t *n_keypoints, limu_frame_stats *stats) {
    LIMU_REQUIRE(o && n >= 0 && stride_bytes >= 12 && stride_bytes % 4 == 0 && (n == 0 || (points && timestamps)), "limu_odom_register_cloud: bad arguments");
    LIMU_TRY(bind(o->ctx));
    LIMU_TRY(stage_in(o->ctx, o->raw, points, (size_t)n * stride_bytes));
    LIMU_TRY(stage_in(o->ctx, o->ts, timestamps, (size_t)n * 8));
    return odom_register_device(o, o->raw.p, 1, stride_bytes, o->ts.as<double>(), n, pose_out, down_xyz, n_down, keypoints_xyz, n_keypoints, stats);
}

int limu_odom_register_frame_dev(limu_odom *o, const float *xyzt_dev, int64_t n, double pose_out[7], limu_frame_stats *stats) {
    LIMU_REQUIRE(o && n >= 0 && (n == 0 || xyzt_dev), "limu_odom_register_frame_dev: bad arguments");
    LIMU_TRY(bind(o->ctx));
    return odom_register_device(o, xyzt_dev, 0, 0, nullptr, n, pose_out, nullptr, nullptr, nullptr, nullptr, stats);
}

int limu_odom_hint_next_dev(limu_odom *o, const float *xyzt_dev_next, int64_t n_next) {
    LIMU_REQUIRE(o && n_next >= 0, "limu_odom_hint_next_dev: bad arguments");
    o->hint_ptr = (o->speculate && n_next > 0) ? xyzt_dev_next : nullptr;   // with LIMU_OPT_SPECULATE off a hint is ignored
    o->hint_n = (o->speculate && n_next > 0) ? n_next : 0;
    return LIMU_OK;
}

int limu_odom_register_points(limu_odom *o, const double *xyz, int64_t n, double pose_out[7], double *down_xyz, int64_t *n_down,
                              double *keypoints_xyz, int64_t *n_keypoints, limu_frame_stats *stats) {
    LIMU_REQUIRE(o && n >= 0 && (n == 0 || xyz), "limu_odom_register_points: bad arguments");
    LIMU_TRY(bind(o->ctx));
    LIMU_TRY(stage_in(o->ctx, o->raw, xyz, (size_t)n * 24));
    return odom_register_device(o, o->raw.p, 2, 0, nullptr, n, pose_out, down_xyz, n_down, keypoints_xyz, n_keypoints, stats);
}

int limu_odom_register_msg(limu_odom *o, const void *data, int64_t n, const limu_cloud_fields *fields, const limu_lidar_config *cfg, double message_time,
                           int32_t scan_count, int32_t max_segments, double *poses_out, int64_t *seg_sizes, double *seg_time, int32_t *n_segments,
                           limu_frame_stats *stats) {
    LIMU_REQUIRE(o && n >= 0 && (n == 0 || data) && n_segments && max_segments >= 0 && (max_segments == 0 || (poses_out && seg_sizes && seg_time)) &&
                     n < (int64_t(1) << 31), "limu_odom_register_msg: bad arguments");
    LIMU_TRY(bind(o->ctx));
    LIMU_TRY(preprocess_validate(fields, cfg, max_segments));
    *n_segments = 0;
    if (n == 0) return LIMU_OK;
    limu_ctx *c = o->ctx;
    LIMU_TRY(stage_in(c, o->pre.raw, data, (size_t)n * fields->point_step));
    LIMU_TRY(preprocess_device(c, o->pre, o->pre.raw.as<unsigned char>(), n, *fields, *cfg, message_time, scan_count));   // lidar_callback, odom_run.cpp:51-67
    const SegTable &T = *o->pre.h_seg;
    const int ns = std::min<int>(T.nseg, max_segments);
    for (int k = 0; k < ns; ++k) {
        const int64_t off = T.begin[k], nk = T.end[k] - T.begin[k];
        seg_sizes[k] = nk; seg_time[k] = T.time[k];
        if (stats) memset(&stats[k], 0, sizeof(limu_frame_stats));
        if (nk <= 1) {   // "Too few input point cloud!" (odom_run.cpp:78-83): the segment is popped without being registered
            pose_store(o->poses.empty() ? pose_identity() : o->poses.back(), poses_out + 7 * k);
            continue;
        }
        // estimate_lidar_odometry -> register_frame(*processed_frame, time_buffer) (odom_run.cpp:103-106) on the device-resident segment
        LIMU_TRY(odom_register_device(o, o->pre.rec.as<unsigned char>() + (size_t)off * 48, 1, 48, o->pre.ts.as<double>() + off, nk, poses_out + 7 * k, nullptr,
                                      nullptr, nullptr, nullptr, stats ? &stats[k] : nullptr));
    }
    *n_segments = ns;
    return LIMU_OK;
}

int limu_odom_num_poses(limu_odom *o, int64_t *n) {
    LIMU_REQUIRE(o && n, "limu_odom_num_poses: null argument");
    *n = (int64_t)o->poses.size();
    return LIMU_OK;
}
int limu_odom_pose(limu_odom *o, int64_t i, double pose_out[7]) {
    LIMU_REQUIRE(o && pose_out && i >= 0 && i < (int64_t)o->poses.size(), "limu_odom_pose: index out of range");
    pose_store(o->poses[(size_t)i], pose_out);
    return LIMU_OK;
}
int limu_odom_adaptive_threshold(limu_odom *o, double *sigma) {
    LIMU_REQUIRE(o && sigma, "limu_odom_adaptive_threshold: null argument");
    *sigma = odom_adaptive_threshold(o);
    return LIMU_OK;
}
int limu_odom_prediction(limu_odom *o, double pose_out[7]) {
    LIMU_REQUIRE(o && pose_out, "limu_odom_prediction: null argument");
    pose_store(odom_prediction(o), pose_out);
    return LIMU_OK;
}
int limu_odom_has_moved(limu_odom *o, int *out) {
    LIMU_REQUIRE(o && out, "limu_odom_has_moved: null argument");
    *out = odom_has_moved(o) ? 1 : 0;
    return LIMU_OK;
}
limu_map *limu_odom_map(limu_odom *o) { return o ? o->map : nullptr; }

int limu_odom_set_option(limu_odom *o, int32_t option, int64_t value) {
    LIMU_REQUIRE(o, "limu_odom_set_option: null handle");
    if (option == LIMU_OPT_SPECULATE) {
        LIMU_TRY(bind(o->ctx));
        if (!value && o->spec_ptr) {   // a launch is in flight for a scan that will now be voxelized again: let it finish first
            LIMU_CUDA_TRY(cudaStreamSynchronize(o->ctx->stream));
            o->spec_ptr = nullptr; o->spec_slot = -1;
        }
        o->speculate = value != 0;
        if (!o->speculate) { o->hint_ptr = nullptr; o->hint_n = 0; }
        return LIMU_OK;
    }
    if (option == LIMU_OPT_CLUSTER_LOOP) { o->cluster_loop = value != 0; return LIMU_OK; }
    set_error("limu_odom_set_option: unknown option %d", (int)option);
    return LIMU_ERR_INVALID;
}

}  // extern "C"
