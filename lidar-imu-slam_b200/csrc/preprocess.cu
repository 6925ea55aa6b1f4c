// preprocess.cu -- the host preprocessing in front of KissICP::register_frame, on the device:
//   frame::Lidar::process_frame   L/src/sensors/lidar/frame.cpp:101-193   range gate, NaN drop, per-point offset time
//   Lidar::sort_clouds            frame.cpp:28-51                         order by offset time (curvature)
//   Lidar::split_clouds           frame.cpp:53-99                         drop the first point, cut into frame_split_num segments
//   utils::get_time_stamps / normalize_timestamps   L/src/utils/calculation_helpers.cpp:3-81
// The raw sensor_msgs::PointCloud2 payload goes to HBM as it is; the outputs are the reference's own processed clouds
// (pcl::PointXYZINormal records, 48 B) + normalised FP64 timestamps, which the odometry kernels read in place
// (limu_odom_register_msg) -- the serial host loop, the index sort and three std::vector copies of the reference disappear.
//
//   k_pre_scan        one thread per message point: gate (FLOAT range arithmetic as frame.cpp:143), curvature, extracted
//                     timestamp, flag; global maximum of the extracted timestamps (normalize_timestamps)
//   k_pre_ring_model  only when the last point carries no offset time (frame.cpp:128-133, :159-182, "constant rotation
//                     model"): one CTA per scan line walks that ring's points in message order; the serial recurrence
//                     curvature_i = f_i(curvature_{i-1}) is a composition of two-valued threshold functions, so a block
//                     scan with a carried state reproduces it exactly
//   compact           order-preserving compaction of the survivors (compact.cuh)
//   k_rs_*            stable LSD radix sort (4 x 8 bit) of (curvature key, source index)
//   k_pre_split       segment table of split_clouds (cut positions, accumulated segment time); k_pre_segmax: per-segment timestamp maximum
//   k_pre_emit        writes the processed records and timestamps
//
// std::sort in sort_clouds is unstable: among EQUAL curvature keys the reference's order is whatever libstdc++'s introsort
// leaves. The device sort is stable (ties keep message order), as the C oracle defines it; results are identical whenever
// the keys are distinct.
#include <algorithm>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "compact.cuh"
#include "ops.cuh"

namespace limu {

constexpr int RS_BLOCK = 512;

struct PreArgs {
    const unsigned char *data;
    int64_t n;
    limu_cloud_fields f;
    double blind_sq, max_sq, angle_limit, scan_ang_vel, message_time;
    int num_scan_lines, cuts;          // cuts = required_cut_num (frame.cpp:64)
    float *curv;                       // per message point
    double *ext;                       // per message point: extracted timestamp (not yet normalised)
    unsigned char *flags;
    unsigned long long *gmax;          // bits of the maximum extracted timestamp (uint32 source: non-negative -> bit order == value order)
    DevStatus *status;
};


template <class T> __device__ __forceinline__ T load_field(const unsigned char *p) {
    T v;
    if ((reinterpret_cast<uintptr_t>(p) & (sizeof(T) - 1)) == 0) return *reinterpret_cast<const T *>(p);
    unsigned char b[sizeof(T)];
#pragma unroll
    for (int k = 0; k < (int)sizeof(T); ++k) b[k] = p[k];
    memcpy(&v, b, sizeof(T));
    return v;
}

__device__ __forceinline__ bool pre_has_offset_time(const PreArgs &A) {   // frame.cpp:128
    if (A.f.off_timestamp < 0 || A.n <= 0) return false;
    return load_field<double>(A.data + (size_t)(A.n - 1) * A.f.point_step + A.f.off_timestamp) > 0;
}

static __global__ void __launch_bounds__(256) k_pre_scan(const PreArgs A) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double ext = 0.0;
    bool on = i < A.n;
    if (on) {
        const unsigned char *r = A.data + (size_t)i * A.f.point_step;
        // utils::get_time_stamps: "time" is read as double; "t"/"timestamp" as uint32 whatever the datatype (calculation_helpers.cpp:31-44)
        ext = A.f.time_field_is_f64 ? load_field<double>(r + A.f.off_time_field) : (double)load_field<unsigned int>(r + A.f.off_time_field);
        A.ext[i] = ext;
        const float x = A.f.off_x >= 0 ? load_field<float>(r + A.f.off_x) : 0.f, y = A.f.off_y >= 0 ? load_field<float>(r + A.f.off_y) : 0.f,
                    z = A.f.off_z >= 0 ? load_field<float>(r + A.f.off_z) : 0.f;
        const float distf = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));   // frame.cpp:143, float arithmetic, no contraction
        const double dist = (double)distf;
        const bool keep = (dist >= A.blind_sq && dist <= A.max_sq) && !(isnan(x) || isnan(y) || isnan(z));   // :144
        const double ts = A.f.off_timestamp >= 0 ? load_field<double>(r + A.f.off_timestamp) : 0.0;
        A.curv[i] = (float)(((ts - A.message_time) + 0.1) * 1000.0);   // :156
        A.flags[i] = keep ? 1 : 0;
    }
    if (!A.f.time_field_is_f64) {   // normalize_timestamps divides by the maximum over ALL message points (:52-66)
        unsigned long long b = on ? (unsigned long long)__double_as_longlong(ext) : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const unsigned long long t = __shfl_xor_sync(0xFFFFFFFFu, b, o); b = t > b ? t : b; }
        if ((threadIdx.x & 31) == 0 && b) atomicMax(A.gmax, b);
    }
}

// ---- constant rotation model (frame.cpp:159-182) ---------------------------------------------------------------
// Per ring, in message order:  first kept point: yaw_fp = yaw, point DROPPED, time_last = 0 (:163-171);
// every later one: base = float(angle_diff / scan_ang_vel); curvature = base < time_last ? float(base + period) : base; time_last = curvature.
// f_i(c) = c > base_i ? hi_i : lo_i is a two-valued threshold function and such functions are closed under composition.
struct StepFn { double T, LO, HI; int id; };
__device__ __forceinline__ double stepfn_apply(const StepFn &f, double c) { return f.id ? c : (c > f.T ? f.HI : f.LO); }
__device__ __forceinline__ StepFn stepfn_then(const StepFn &a, const StepFn &b) {   // first a, then b
    if (a.id) return b;
    if (b.id) return a;
    StepFn r; r.id = 0; r.T = a.T; r.LO = stepfn_apply(b, a.LO); r.HI = stepfn_apply(b, a.HI);
    return r;
}
__device__ __forceinline__ StepFn stepfn_shfl_up(const StepFn &f, int o) {
    StepFn r;
    r.T = __shfl_up_sync(0xFFFFFFFFu, f.T, o); r.LO = __shfl_up_sync(0xFFFFFFFFu, f.LO, o); r.HI = __shfl_up_sync(0xFFFFFFFFu, f.HI, o);
    r.id = __shfl_up_sync(0xFFFFFFFFu, f.id, o);
    return r;
}

static __global__ void __launch_bounds__(1024) k_pre_ring_model(const PreArgs A) {
    if (pre_has_offset_time(A)) return;
    __shared__ StepFn wtot[32];
    __shared__ int first_idx;
    __shared__ double s_yaw_fp, s_last;
    __shared__ int s_have_first;
    const int ring = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { s_have_first = 0; s_yaw_fp = 0.0; s_last = 0.0; }
    const double period = A.angle_limit / A.scan_ang_vel;
    __syncthreads();
    for (int64_t base = 0; base < A.n; base += 1024) {
        const int64_t i = base + threadIdx.x;
        bool mine = false;
        double yaw = 0.0;
        if (i < A.n && A.flags[i]) {
            const unsigned char *r = A.data + (size_t)i * A.f.point_step;
            const int layer = A.f.off_ring >= 0 ? (int)load_field<unsigned short>(r + A.f.off_ring) : 0;
            if (layer >= A.num_scan_lines) { if (ring == 0) A.status->pad[0] = 1; }   // the reference indexes its per-ring vectors out of bounds
            else if (layer == ring) {
                mine = true;
                const float x = A.f.off_x >= 0 ? load_field<float>(r + A.f.off_x) : 0.f, y = A.f.off_y >= 0 ? load_field<float>(r + A.f.off_y) : 0.f;
                yaw = (double)(float)atan2((double)y, (double)x) * 57.2957;   // atan2f of the float members (:161)
            }
        }
        // the ring's first kept point of the whole message: sets yaw_fp and is dropped
        const int had_first = s_have_first;
        if (threadIdx.x == 0) first_idx = 0x7FFFFFFF;
        __syncthreads();
        if (!had_first && mine) atomicMin(&first_idx, (int)threadIdx.x);
        __syncthreads();
        const bool is_first = !had_first && mine && first_idx == (int)threadIdx.x;
        if (is_first) { s_yaw_fp = yaw; s_have_first = 1; A.flags[i] = 0; }
        __syncthreads();
        const double yaw_fp = s_yaw_fp;
        StepFn f; f.id = 1; f.T = f.LO = f.HI = 0.0;
        if (mine && !is_first) {
            const double angle_diff = yaw <= yaw_fp ? (yaw_fp - yaw) : ((yaw_fp - yaw) + A.angle_limit);   // :174
            const float b = (float)(angle_diff / A.scan_ang_vel);                                            // :175
            f.id = 0; f.T = (double)b; f.LO = (double)b; f.HI = (double)(float)((double)b + period);         // :177-178
        }
        // inclusive block scan under composition (earlier elements first)
        StepFn x = f;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const StepFn t = stepfn_shfl_up(x, o); if (lane >= o) x = stepfn_then(t, x); }
        if (lane == 31) wtot[warp] = x;
        __syncthreads();
        if (warp == 0) {
            StepFn w = wtot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const StepFn t = stepfn_shfl_up(w, o); if (lane >= o) w = stepfn_then(t, w); }
            wtot[lane] = w;   // inclusive over warps
        }
        __syncthreads();
        if (warp > 0) x = stepfn_then(wtot[warp - 1], x);
        const double last = s_last;
        if (mine && !is_first) A.curv[i] = (float)stepfn_apply(x, last);
        __syncthreads();
        if (threadIdx.x == 1023) s_last = stepfn_apply(x, last);   // carried time_last (:181)
        __syncthreads();
    }
}

// ---- stable LSD radix sort of (u32 key, u32 value), n read from device memory ------------------------------------
__device__ __forceinline__ unsigned int float_key(float v) {
    const unsigned int u = __float_as_uint(v);
    return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
static __global__ void __launch_bounds__(256) k_pre_keys(const float *__restrict__ curv, const int *__restrict__ sidx, const int *m_dev,
                                                        unsigned int *__restrict__ keys, unsigned int *__restrict__ vals) {
    const int m = *m_dev;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const int i = sidx[j];
    keys[j] = float_key(curv[i]);
    vals[j] = (unsigned int)i;
}
static __global__ void __launch_bounds__(RS_BLOCK) k_rs_hist(const unsigned int *__restrict__ keys, const int *m_dev, int shift, int *__restrict__ hist, int B) {
    __shared__ int sh[256];
    const int m = *m_dev;
    if (threadIdx.x < 256) sh[threadIdx.x] = 0;
    __syncthreads();
    const int j = blockIdx.x * RS_BLOCK + threadIdx.x;
    if (j < m) atomicAdd(&sh[(keys[j] >> shift) & 0xFF], 1);
    __syncthreads();
    if (threadIdx.x < 256) hist[threadIdx.x * B + blockIdx.x] = sh[threadIdx.x];
}
// hist is digit-major [256][B]. One CTA per digit row: exclusive scan of the row in place, row total -> totals[digit].
// (A single-CTA scan of all 256*B counters cost 32 us per pass at 128k points; the rows are independent.)
static __global__ void __launch_bounds__(256) k_rs_rowscan(int *__restrict__ hist, int B, int *__restrict__ totals) {
    __shared__ int ws[8];
    __shared__ int carry;
    int *row = hist + (size_t)blockIdx.x * B;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < B; base += 256) {
        const int at = base + threadIdx.x;
        const int v = at < B ? row[at] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) ws[warp] = incl;
        __syncthreads();
        int wbase = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) wbase += w < warp ? ws[w] : 0;
        const int c = carry;
        if (at < B) row[at] = c + wbase + incl - v;
        __syncthreads();
        if (threadIdx.x == 255) carry = c + wbase + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) totals[blockIdx.x] = carry;
}
static __global__ void __launch_bounds__(RS_BLOCK) k_rs_scatter(const unsigned int *__restrict__ keys, const unsigned int *__restrict__ vals, const int *m_dev,
                                                               int shift, const int *__restrict__ hist, const int *__restrict__ totals, int B,
                                                               unsigned int *__restrict__ keys_out, unsigned int *__restrict__ vals_out) {
    __shared__ int cnt[RS_BLOCK / 32][256];
    __shared__ int dbase[256];   // exclusive scan of the 256 digit totals: where each digit's range starts in the output
    __shared__ int dws[8];
    const int m = *m_dev;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = threadIdx.x; k < (RS_BLOCK / 32) * 256; k += RS_BLOCK) (&cnt[0][0])[k] = 0;
    __syncthreads();
    const int j = blockIdx.x * RS_BLOCK + threadIdx.x;
    const bool on = j < m;
    const unsigned int key = on ? keys[j] : 0u, val = on ? vals[j] : 0u;
    const int digit = on ? (int)((key >> shift) & 0xFF) : 256;
    const unsigned int peers = __match_any_sync(0xFFFFFFFFu, digit);
    const int rank_in_warp = __popc(peers & ((1u << lane) - 1u));
    if (on && rank_in_warp == 0) cnt[warp][digit] = __popc(peers);
    __syncthreads();
    if (threadIdx.x < 256) {   // warps 0..7 scan the digit totals
        const int v = totals[threadIdx.x];
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) dws[warp] = incl;
        dbase[threadIdx.x] = incl - v;
    }
    __syncthreads();
    if (threadIdx.x < 256) {
        int run = dbase[threadIdx.x] + hist[threadIdx.x * B + blockIdx.x];
#pragma unroll
        for (int w = 0; w < 8; ++w) run += w < warp ? dws[w] : 0;
#pragma unroll
        for (int w = 0; w < RS_BLOCK / 32; ++w) { const int c = cnt[w][threadIdx.x]; cnt[w][threadIdx.x] = run; run += c; }
    }
    __syncthreads();
    if (on) {
        const int pos = cnt[warp][digit] + rank_in_warp;
        keys_out[pos] = key;
        vals_out[pos] = val;
    }
}

// ---- split_clouds (frame.cpp:53-99) ------------------------------------------------------------------------------
// Sorted position 0 is never emitted (the loop starts at 1); segment k ends at the position where the running count equals
// (int)((k+1) * m / cuts - 1) -- evaluated in size_t like the reference -- and a threshold that is not ahead of the running
// count is never met, which ends the cutting. One CTA: thread 0 lays out the table, then all threads reduce each segment's
// timestamp maximum.
__device__ __forceinline__ double pre_normalised_ts(const double *ext, unsigned int i, bool global_norm, double gmax) {
    const double v = ext[i];
    return global_norm ? v / gmax : v;
}
// order-preserving map double -> u64 (for atomicMax over possibly negative timestamps) and back
__device__ __forceinline__ unsigned long long f64_ordered(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double f64_unordered(unsigned long long k) {
    return __longlong_as_double((long long)((k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k));
}
static __global__ void k_pre_split(const PreArgs A, const int *m_dev, const unsigned int *__restrict__ vals, SegTable *T, unsigned long long *tmax_keys) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int m = *m_dev;
    const double message_time_ms = A.message_time * 1000;
    double last_end = message_time_ms;
    int nseg = 0, prev = 0;
    const unsigned long long M = (unsigned long long)(m > 0 ? m : 0), cuts = (unsigned long long)A.cuts;
    for (int cut = 0; cut < PRE_MAX_SEG && m > 1; ++cut) {
        const int thr = (int)(((unsigned long long)(cut + 1) * M / cuts) - 1ull);
        if (thr <= prev || thr > m - 1) break;
        const double adj = message_time_ms - last_end;
        T->begin[nseg] = prev; T->end[nseg] = thr; T->adj[nseg] = adj; T->time[nseg] = last_end / (double)1000;
        tmax_keys[nseg] = 0ull;                                          // below every ordered key
        const float c_end = (float)((double)A.curv[vals[thr]] + adj);   // :74 on the cut point
        last_end += (double)c_end;                                    // :92
        prev = thr;
        ++nseg;
    }
    T->nseg = nseg; T->m = m;
}
// per-segment timestamp maximum (normalize_timestamps :87): grid-stride over the sorted positions, warp-reduced per segment
static __global__ void __launch_bounds__(256) k_pre_segmax(const PreArgs A, const unsigned int *__restrict__ vals, const SegTable *T, unsigned long long *tmax_keys) {
    __shared__ int s_end[PRE_MAX_SEG];
    const int nseg = T->nseg;
    for (int k = threadIdx.x; k < nseg; k += blockDim.x) s_end[k] = T->end[k];
    __syncthreads();
    if (nseg == 0) return;
    const int last = s_end[nseg - 1];
    const double gmax = __longlong_as_double((long long)*A.gmax);
    const bool global_norm = !A.f.time_field_is_f64 && !(gmax < 1.0);
    for (int p0 = 1 + blockIdx.x * blockDim.x; p0 <= last; p0 += gridDim.x * blockDim.x) {   // whole warps iterate together
        const int p = p0 + threadIdx.x;
        int k = -1;
        unsigned long long key = 0ull;
        if (p <= last) {
            k = 0;
            while (p > s_end[k]) ++k;
            key = f64_ordered(pre_normalised_ts(A.ext, vals[p], global_norm, gmax));
        }
        // lanes of a warp almost always share the segment: reduce within the warp when they do
        const int k0 = __shfl_sync(0xFFFFFFFFu, k, 0);
        if (__all_sync(0xFFFFFFFFu, k == k0)) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { const unsigned long long t = __shfl_xor_sync(0xFFFFFFFFu, key, o); key = t > key ? t : key; }
            if ((threadIdx.x & 31) == 0 && k >= 0) atomicMax(tmax_keys + k, key);
        } else if (k >= 0) {
            atomicMax(tmax_keys + k, key);
        }
    }
}
static __global__ void k_pre_segmax_store(SegTable *T, const unsigned long long *tmax_keys) {
    const int k = threadIdx.x;
    if (k < T->nseg) T->tmax[k] = f64_unordered(tmax_keys[k]);
}

// Output rows: pcl::PointXYZINormal (48 B): {x, y, z, 1}, {normal 0,0,0, 0}, {intensity, curvature, 0, 0}; + FP64 timestamp.
static __global__ void __launch_bounds__(256) k_pre_emit(const PreArgs A, const unsigned int *__restrict__ vals, const SegTable *T, float4 *__restrict__ out_rec,
                                                        double *__restrict__ out_ts) {
    __shared__ int s_end[PRE_MAX_SEG];
    __shared__ double s_adj[PRE_MAX_SEG], s_tmax[PRE_MAX_SEG];
    const int nseg = T->nseg;
    for (int k = threadIdx.x; k < nseg; k += blockDim.x) { s_end[k] = T->end[k]; s_adj[k] = T->adj[k]; s_tmax[k] = T->tmax[k]; }
    __syncthreads();
    if (nseg == 0) return;
    const int last = s_end[nseg - 1];
    const double gmax = __longlong_as_double((long long)*A.gmax);
    const bool global_norm = !A.f.time_field_is_f64 && !(gmax < 1.0);
    for (int p = 1 + blockIdx.x * blockDim.x + threadIdx.x; p <= last; p += gridDim.x * blockDim.x) {
        int k = 0;
        while (p > s_end[k]) ++k;
        const unsigned int i = vals[p];
        const unsigned char *r = A.data + (size_t)i * A.f.point_step;
        const float x = A.f.off_x >= 0 ? load_field<float>(r + A.f.off_x) : 0.f, y = A.f.off_y >= 0 ? load_field<float>(r + A.f.off_y) : 0.f,
                    z = A.f.off_z >= 0 ? load_field<float>(r + A.f.off_z) : 0.f;
        const float inten = A.f.off_intensity >= 0 ? (float)r[A.f.off_intensity] : 0.f;   // :154 uint8 -> float
        const float c = (float)((double)A.curv[i] + s_adj[k]);                             // :74
        const double t = pre_normalised_ts(A.ext, i, global_norm, gmax);
        float4 *o = out_rec + 3 * (size_t)(p - 1);
        o[0] = make_float4(x, y, z, 1.0f);
        o[1] = make_float4(0.f, 0.f, 0.f, 0.f);
        o[2] = make_float4(inten, c, 0.f, 0.f);
        out_ts[p - 1] = s_tmax[k] < 1.0 ? t : t / s_tmax[k];                              // normalize_timestamps per segment :87
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Everything up to the processed records in device memory. data_dev: the message payload on the device.
// On return sc.rec / sc.ts hold the concatenated segments and *sc.h_seg the segment table (host, after one sync).
int preprocess_device(limu_ctx *c, PreScratch &sc, const unsigned char *data_dev, int64_t n, const limu_cloud_fields &f, const limu_lidar_config &cfg,
                      double message_time, int scan_count) {
    if (!sc.h_seg) LIMU_CUDA_TRY(cudaHostAlloc(&sc.h_seg, sizeof(SegTable), cudaHostAllocDefault));
    memset(sc.h_seg, 0, sizeof(SegTable));
    if (n <= 0) return LIMU_OK;
    const size_t N = (size_t)n;
    LIMU_TRY(sc.curv.reserve(N * 4, c->stream));
    LIMU_TRY(sc.ext.reserve(N * 8, c->stream));
    LIMU_TRY(sc.flags.reserve(N, c->stream));
    LIMU_TRY(sc.blockcnt.reserve((size_t)div_up(n, COMPACT_BLOCK) * 4 + 16, c->stream));
    LIMU_TRY(sc.sidx.reserve(N * 4 + 16, c->stream));
    for (int k = 0; k < 2; ++k) { LIMU_TRY(sc.keys[k].reserve(N * 4, c->stream)); LIMU_TRY(sc.vals[k].reserve(N * 4, c->stream)); }
    const int B = div_up(n, RS_BLOCK);
    LIMU_TRY(sc.hist.reserve((size_t)B * 256 * 4 + 256 * 4, c->stream));   // [256][B] block counts + 256 digit totals
    LIMU_TRY(sc.small.reserve(256 + sizeof(SegTable) + PRE_MAX_SEG * 8, c->stream));
    LIMU_TRY(sc.rec.reserve(N * 48, c->stream));
    LIMU_TRY(sc.ts.reserve(N * 8, c->stream));
    unsigned long long *gmax = sc.small.as<unsigned long long>();
    int *m_dev = reinterpret_cast<int *>(gmax + 1);
    SegTable *seg_dev = reinterpret_cast<SegTable *>(sc.small.as<unsigned char>() + 256);
    unsigned long long *tmax_keys = reinterpret_cast<unsigned long long *>(sc.small.as<unsigned char>() + 256 + sizeof(SegTable));
    int *totals = sc.hist.as<int>() + (size_t)B * 256;
    LIMU_CUDA_TRY(cudaMemsetAsync(sc.small.p, 0, 256, c->stream));

    PreArgs A;
    memset(&A, 0, sizeof A);
    A.data = data_dev; A.n = n; A.f = f;
    A.blind_sq = cfg.min_range * cfg.min_range; A.max_sq = cfg.max_range * cfg.max_range; A.angle_limit = cfg.max_angle - cfg.min_angle;   // lidar/frame.hpp:141-146
    A.scan_ang_vel = (double)(int)cfg.frame_rate * (360.0 / 1000.0);   // utils::calc_scan_ang_vel(int), calculation_helpers.cpp:104-108
    A.message_time = message_time;
    A.num_scan_lines = cfg.num_scan_lines;
    A.cuts = scan_count < 20 ? 1 : cfg.frame_split_num;   // MIN_SCAN_COUNT, frame.cpp:5,:64
    A.curv = sc.curv.as<float>(); A.ext = sc.ext.as<double>(); A.flags = sc.flags.as<unsigned char>(); A.gmax = gmax; A.status = c->d_status;

    k_pre_scan<<<div_up(n, 256), 256, 0, c->stream>>>(A);
    LIMU_LAUNCHED();
    k_pre_ring_model<<<std::max(1, cfg.num_scan_lines), 1024, 0, c->stream>>>(A);   // returns at once when the scan carries offset times
    LIMU_LAUNCHED();
    LIMU_TRY(compact_flags(c, A.flags, n, nullptr, sc.blockcnt.as<int>(), sc.sidx.as<int>(), m_dev));
    k_pre_keys<<<div_up(n, 256), 256, 0, c->stream>>>(A.curv, sc.sidx.as<int>(), m_dev, sc.keys[0].as<unsigned int>(), sc.vals[0].as<unsigned int>());
    LIMU_LAUNCHED();
    int cur = 0;
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 8 * pass;
        k_rs_hist<<<B, RS_BLOCK, 0, c->stream>>>(sc.keys[cur].as<unsigned int>(), m_dev, shift, sc.hist.as<int>(), B);
        LIMU_LAUNCHED();
        k_rs_rowscan<<<256, 256, 0, c->stream>>>(sc.hist.as<int>(), B, totals);
        LIMU_LAUNCHED();
        k_rs_scatter<<<B, RS_BLOCK, 0, c->stream>>>(sc.keys[cur].as<unsigned int>(), sc.vals[cur].as<unsigned int>(), m_dev, shift, sc.hist.as<int>(), totals, B,
                                                    sc.keys[cur ^ 1].as<unsigned int>(), sc.vals[cur ^ 1].as<unsigned int>());
        LIMU_LAUNCHED();
        cur ^= 1;
    }
    k_pre_split<<<1, 32, 0, c->stream>>>(A, m_dev, sc.vals[cur].as<unsigned int>(), seg_dev, tmax_keys);
    LIMU_LAUNCHED();
    const int eb = (int)std::min<int64_t>(div_up(n, 256), (int64_t)c->sm_count * 8);
    k_pre_segmax<<<eb, 256, 0, c->stream>>>(A, sc.vals[cur].as<unsigned int>(), seg_dev, tmax_keys);
    LIMU_LAUNCHED();
    k_pre_segmax_store<<<1, PRE_MAX_SEG, 0, c->stream>>>(seg_dev, tmax_keys);
    LIMU_LAUNCHED();
    k_pre_emit<<<eb, 256, 0, c->stream>>>(A, sc.vals[cur].as<unsigned int>(), seg_dev, sc.rec.as<float4>(), sc.ts.as<double>());
    LIMU_LAUNCHED();
    LIMU_CUDA_TRY(cudaMemcpyAsync(sc.h_seg, seg_dev, sizeof(SegTable), cudaMemcpyDeviceToHost, c->stream));
    LIMU_TRY(check_status(c));   // synchronises
    return LIMU_OK;
}

static std::mutex g_pre_mu;
static std::unordered_map<limu_ctx *, PreScratch *> g_pre;
PreScratch *pre_scratch_of(limu_ctx *c) {
    std::lock_guard<std::mutex> lk(g_pre_mu);
    auto it = g_pre.find(c);
    if (it == g_pre.end()) it = g_pre.emplace(c, new PreScratch).first;
    return it->second;
}
void release_pre_scratch(limu_ctx *c) {
    std::lock_guard<std::mutex> lk(g_pre_mu);
    auto it = g_pre.find(c);
    if (it == g_pre.end()) return;
    it->second->release();
    delete it->second;
    g_pre.erase(it);
}

int preprocess_validate(const limu_cloud_fields *f, const limu_lidar_config *cfg, int max_segments) {
    LIMU_REQUIRE(f && cfg, "limu_preprocess_frame: null fields / config");
    LIMU_REQUIRE(f->point_step > 0 && f->off_time_field >= 0 && f->off_time_field + (f->time_field_is_f64 ? 8 : 4) <= f->point_step,
                 "limu_preprocess_frame: the message needs a 't', 'timestamp' or 'time' field inside each point record");
    const int offs[6] = {f->off_x, f->off_y, f->off_z, f->off_intensity, f->off_ring, f->off_timestamp}, sz[6] = {4, 4, 4, 1, 2, 8};
    for (int k = 0; k < 6; ++k) LIMU_REQUIRE(offs[k] < 0 || offs[k] + sz[k] <= f->point_step, "limu_preprocess_frame: a field lies outside the point record");
    LIMU_REQUIRE(cfg->frame_split_num >= 1 && cfg->frame_split_num <= PRE_MAX_SEG && cfg->num_scan_lines >= 1 && cfg->num_scan_lines <= 4096,
                 "limu_preprocess_frame: need 1 <= frame_split_num <= 64 and 1 <= num_scan_lines <= 4096");
    LIMU_REQUIRE(max_segments >= 0, "limu_preprocess_frame: negative max_segments");
    return LIMU_OK;
}

}  // namespace limu

using namespace limu;

extern "C" {

void limu_lidar_default_config(limu_lidar_config *cfg) {   // lidar/frame.hpp:64-70
    if (!cfg) return;
    cfg->min_range = 5.0; cfg->max_range = 100.0; cfg->min_angle = 0.0; cfg->max_angle = 360.0; cfg->frame_rate = 10.0;
    cfg->num_scan_lines = 16; cfg->frame_split_num = 1;
}

int limu_cloud_fields_from_pointfields(int32_t nf, const char *names, const int32_t *offsets, const int32_t *datatypes, const int32_t *counts,
                                       int32_t point_step, limu_cloud_fields *out) {
    LIMU_REQUIRE(out && nf >= 0 && (nf == 0 || (names && offsets && datatypes && counts)) && point_step > 0, "limu_cloud_fields_from_pointfields: bad arguments");
    enum { F_UINT8 = 2, F_UINT16 = 4, F_FLOAT32 = 7, F_FLOAT64 = 8 };   // sensor_msgs::PointField
    out->point_step = point_step;
    out->off_x = out->off_y = out->off_z = out->off_intensity = out->off_ring = out->off_timestamp = out->off_time_field = -1;
    out->time_field_is_f64 = 0;
    int tsf_count = 0;
    const char *p = names;
    for (int k = 0; k < nf; ++k) {
        const size_t len = strlen(p);
        // pcl::fromROSMsg: a LidarPoint member is filled from the first field of the same name AND datatype (lidar/frame.hpp:21-23)
        if (!strcmp(p, "x") && datatypes[k] == F_FLOAT32 && out->off_x < 0) out->off_x = offsets[k];
        if (!strcmp(p, "y") && datatypes[k] == F_FLOAT32 && out->off_y < 0) out->off_y = offsets[k];
        if (!strcmp(p, "z") && datatypes[k] == F_FLOAT32 && out->off_z < 0) out->off_z = offsets[k];
        if (!strcmp(p, "intensity") && datatypes[k] == F_UINT8 && out->off_intensity < 0) out->off_intensity = offsets[k];
        if (!strcmp(p, "ring") && datatypes[k] == F_UINT16 && out->off_ring < 0) out->off_ring = offsets[k];
        if (!strcmp(p, "timestamp") && datatypes[k] == F_FLOAT64 && out->off_timestamp < 0) out->off_timestamp = offsets[k];
        // utils::get_time_stamps: the LAST field named t / timestamp / time wins (calculation_helpers.cpp:5-19)
        if (!strcmp(p, "t") || !strcmp(p, "timestamp") || !strcmp(p, "time")) {
            out->off_time_field = offsets[k]; out->time_field_is_f64 = strcmp(p, "time") == 0 ? 1 : 0; tsf_count = counts[k];
        }
        p += len + 1;
    }
    if (out->off_time_field < 0 || tsf_count == 0) {
        out->off_time_field = -1;
        set_error("Field 't', 'timestamp' or 'time' not existing");   // the reference throws std::runtime_error with this text
        return LIMU_ERR_INVALID;
    }
    return LIMU_OK;
}

int limu_preprocess_frame(limu_ctx *c, const void *data, int64_t n, const limu_cloud_fields *fields, const limu_lidar_config *cfg, double message_time,
                          int32_t scan_count, void *out_points, double *out_ts, int32_t max_segments, int64_t *seg_sizes, double *seg_time, int32_t *n_segments) {
    LIMU_TRY(bind(c));
    LIMU_TRY(preprocess_validate(fields, cfg, max_segments));
    LIMU_REQUIRE(n >= 0 && (n == 0 || data) && n_segments && (max_segments == 0 || (seg_sizes && seg_time)) && n < (int64_t(1) << 31),
                 "limu_preprocess_frame: bad arguments");
    *n_segments = 0;
    if (n == 0) return LIMU_OK;
    PreScratch &sc = *pre_scratch_of(c);
    LIMU_TRY(stage_in(c, sc.raw, data, (size_t)n * fields->point_step));
    LIMU_TRY(preprocess_device(c, sc, sc.raw.as<unsigned char>(), n, *fields, *cfg, message_time, scan_count));
    const SegTable &T = *sc.h_seg;
    const int ns = std::min<int>(T.nseg, max_segments);
    int64_t total = 0;
    for (int k = 0; k < ns; ++k) { seg_sizes[k] = T.end[k] - T.begin[k]; seg_time[k] = T.time[k]; total += seg_sizes[k]; }
    *n_segments = ns;
    if (total > 0 && out_points) LIMU_CUDA_TRY(cudaMemcpyAsync(out_points, sc.rec.p, (size_t)total * 48, cudaMemcpyDeviceToHost, c->stream));
    if (total > 0 && out_ts) LIMU_CUDA_TRY(cudaMemcpyAsync(out_ts, sc.ts.p, (size_t)total * 8, cudaMemcpyDeviceToHost, c->stream));
    LIMU_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return LIMU_OK;
}

}  // extern "C"
