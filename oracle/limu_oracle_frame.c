/* limu_oracle_frame.c -- ORACLE (test infrastructure, not product code): plain-C restatement of the reference's host
 * preprocessing frame::Lidar::process_frame (L/src/sensors/lidar/frame.cpp:101-193) with sort_clouds (:28-51),
 * split_clouds (:53-99) and the timestamp extraction it calls (utils::get_time_stamps / normalize_timestamps,
 * L/src/utils/calculation_helpers.cpp:3-81). SURVEY section 8(f) N3. Pinned against the compiled reference
 * (oracle/_ref, ref_frame_driver.cpp) by tests/test_oracle_pin.py.
 *
 * Two behaviours of the reference's container are DEFINED here (the device path follows the same definitions):
 *   - sort_clouds uses std::sort (unstable): the relative order of points with EQUAL curvature is whatever libstdc++'s
 *     introsort leaves. This restatement sorts stably (ties keep message order). Results are identical whenever the
 *     curvature keys are distinct; with ties they agree as multisets inside every run of equal keys.
 *   - pcl::fromROSMsg fills a LidarPoint member only from a field of the same name AND datatype (PCL field map);
 *     members without such a field stay 0.
 * Floating point: the range test is FLOAT arithmetic ((x*x + y*y) + z*z, frame.cpp:143), curvature is FLOAT
 * (pcl::PointXYZINormal), atan2 on float members is atan2f (libstdc++ <math.h> overloads); built -ffp-contract=off. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "limu_oracle.h"

enum { PF_INT8 = 1, PF_UINT8, PF_INT16, PF_UINT16, PF_INT32, PF_UINT32, PF_FLOAT32, PF_FLOAT64 };   /* sensor_msgs::PointField */

static int find_field(int nf, const char *names, const int *datatypes, const char *want, int datatype) {
    const char *p = names;
    for (int k = 0; k < nf; ++k) {
        if (strcmp(p, want) == 0 && (datatype < 0 || datatypes[k] == datatype)) return k;
        p += strlen(p) + 1;
    }
    return -1;
}

/* stable merge sort of indices by float key (ascending) */
static void merge_sort_idx(int *idx, int *tmp, const float *key, long lo, long hi) {
    if (hi - lo < 2) return;
    const long mid = lo + (hi - lo) / 2;
    merge_sort_idx(idx, tmp, key, lo, mid);
    merge_sort_idx(idx, tmp, key, mid, hi);
    long a = lo, b = mid, o = lo;
    while (a < mid && b < hi) tmp[o++] = (key[idx[b]] < key[idx[a]]) ? idx[b++] : idx[a++];
    while (a < mid) tmp[o++] = idx[a++];
    while (b < hi) tmp[o++] = idx[b++];
    memcpy(idx + lo, tmp + lo, (size_t)(hi - lo) * sizeof(int));
}

long lo_process_frame(const unsigned char *data, long n, int point_step, int nf, const char *names, const int *offsets, const int *datatypes,
                      const int *counts, const double *cfg, double message_time, int scan_count, long max_segments, long *seg_sizes, double *seg_time,
                      float *rec5, double *ts_out) {
    const double min_range = cfg[0], max_range = cfg[1], min_angle = cfg[2], max_angle = cfg[3], frame_rate = cfg[4];
    const int num_scan_lines = (int)cfg[5], frame_split_num = (int)cfg[6];
    const double blind_sq = min_range * min_range, max_sq = max_range * max_range, angle_limit = max_angle - min_angle;   /* frame.hpp:141-146 */
    const double scan_ang_vel = (double)(int)frame_rate * (360.0 / 1000.0);   /* calc_scan_ang_vel(int), calculation_helpers.cpp:104-108 */
    if (n <= 0 || frame_split_num < 1) return 0;

    /* utils::get_time_stamps: the LAST field named t / timestamp / time (calculation_helpers.cpp:5-19) */
    int tsf = -1;
    {
        const char *p = names;
        for (int k = 0; k < nf; ++k) {
            if (!strcmp(p, "t") || !strcmp(p, "timestamp") || !strcmp(p, "time")) tsf = k;
            p += strlen(p) + 1;
        }
    }
    if (tsf < 0 || counts[tsf] == 0) return -1;   /* std::runtime_error("Field 't', 'timestamp' or 'time' not existing") */
    const char *tsname = names;
    for (int k = 0; k < tsf; ++k) tsname += strlen(tsname) + 1;
    const int ts_is_time = strcmp(tsname, "time") == 0;
    double *extracted = (double *)malloc((size_t)n * sizeof(double));
    double mx = 0.0;
    for (long i = 0; i < n; ++i) {
        const unsigned char *q = data + (size_t)i * point_step + offsets[tsf];
        if (ts_is_time) { double v; memcpy(&v, q, 8); extracted[i] = v; }                 /* PointCloud2ConstIterator<double> :41-44 */
        else { uint32_t v; memcpy(&v, q, 4); extracted[i] = (double)v; }                 /* PointCloud2ConstIterator<uint32_t> :31-35, whatever the datatype */
        if (i == 0 || extracted[i] > mx) mx = extracted[i];
    }
    if (!ts_is_time && !(mx < 1.0))                                                       /* normalize_timestamps :52-66 */
        for (long i = 0; i < n; ++i) extracted[i] = extracted[i] / mx;

    /* pcl::fromROSMsg field map of LidarPoint (lidar/frame.hpp:12-23) */
    const int fx = find_field(nf, names, datatypes, "x", PF_FLOAT32), fy = find_field(nf, names, datatypes, "y", PF_FLOAT32),
              fz = find_field(nf, names, datatypes, "z", PF_FLOAT32), fi = find_field(nf, names, datatypes, "intensity", PF_UINT8),
              fr = find_field(nf, names, datatypes, "ring", PF_UINT16), ft = find_field(nf, names, datatypes, "timestamp", PF_FLOAT64);
#define LOADF(f, i, dst) do { if ((f) >= 0) memcpy(&(dst), data + (size_t)(i) * point_step + offsets[f], sizeof(dst)); } while (0)
    double last_ts = 0.0;
    LOADF(ft, n - 1, last_ts);
    const int has_offset_time = last_ts > 0;                                              /* frame.cpp:128 */

    float *px = (float *)malloc((size_t)n * 5 * sizeof(float));   /* survivors: x y z intensity curvature */
    double *vt = (double *)malloc((size_t)n * sizeof(double));
    int *is_first = (int *)calloc((size_t)(num_scan_lines > 0 ? num_scan_lines : 1), sizeof(int));
    double *yaw_fp = (double *)calloc((size_t)(num_scan_lines > 0 ? num_scan_lines : 1), sizeof(double));
    double *time_last = (double *)calloc((size_t)(num_scan_lines > 0 ? num_scan_lines : 1), sizeof(double));
    if (!has_offset_time) for (int k = 0; k < num_scan_lines; ++k) is_first[k] = 1;   /* :129-133 */
    long m = 0;
    long status = 0;
    for (long i = 0; i < n; ++i) {
        float x = 0.f, y = 0.f, z = 0.f;
        uint8_t inten = 0;
        uint16_t ring = 0;
        double tstamp = 0.0;
        LOADF(fx, i, x); LOADF(fy, i, y); LOADF(fz, i, z); LOADF(fi, i, inten); LOADF(fr, i, ring); LOADF(ft, i, tstamp);
        const float distf = (x * x + y * y) + z * z;                                      /* :143 float arithmetic */
        const double dist = (double)distf;
        if (!(dist >= blind_sq && dist <= max_sq) || isnan(x) || isnan(y) || isnan(z)) continue;   /* :144 */
        float curvature = (float)(((tstamp - message_time) + 0.1) * 1000.0);               /* :156 */
        if (!has_offset_time) {
            const int layer = (int)ring;
            if (layer >= num_scan_lines) { status = -2; break; }   /* the reference indexes its per-ring vectors out of bounds here */
            const double yaw_angle = (double)atan2f(y, x) * 57.2957;                       /* :161 */
            if (is_first[layer]) { yaw_fp[layer] = yaw_angle; is_first[layer] = 0; time_last[layer] = 0.0; continue; }   /* :163-171 */
            const double angle_diff = yaw_angle <= yaw_fp[layer] ? (yaw_fp[layer] - yaw_angle) : ((yaw_fp[layer] - yaw_angle) + angle_limit);   /* :174 */
            curvature = (float)(angle_diff / scan_ang_vel);                                /* :175 */
            if ((double)curvature < time_last[layer]) curvature = (float)((double)curvature + angle_limit / scan_ang_vel);   /* :177-178 */
            time_last[layer] = (double)curvature;                                          /* :181 */
        }
        px[5 * m] = x; px[5 * m + 1] = y; px[5 * m + 2] = z; px[5 * m + 3] = (float)inten; px[5 * m + 4] = curvature;
        vt[m] = extracted[i];                                                              /* :185 */
        ++m;
    }
#undef LOADF
    long nseg = 0;
    if (status == 0 && m > 0) {
        /* sort_clouds :28-51 (stable; see header) */
        int *idx = (int *)malloc((size_t)m * sizeof(int)), *tmp = (int *)malloc((size_t)m * sizeof(int));
        float *key = (float *)malloc((size_t)m * sizeof(float));
        for (long j = 0; j < m; ++j) { idx[j] = (int)j; key[j] = px[5 * j + 4]; }
        merge_sort_idx(idx, tmp, key, 0, m);
        /* split_clouds :53-99 */
        const double message_time_ms = message_time * 1000;
        double last_frame_end_time = message_time_ms;
        const size_t valid_pcl_size = (size_t)m;
        int valid_num = 0, cut_num = 0;
        const int required_cut_num = (scan_count < 20) ? 1 : frame_split_num;             /* MIN_SCAN_COUNT frame.cpp:5, :64 */
        long at = 0, seg_begin = 0;
        for (long id = 1; id < m; ++id) {
            valid_num++;
            const int s = idx[id];
            const float c = (float)((double)px[5 * s + 4] + (message_time_ms - last_frame_end_time));   /* :74 float += double */
            rec5[5 * at] = px[5 * s]; rec5[5 * at + 1] = px[5 * s + 1]; rec5[5 * at + 2] = px[5 * s + 2]; rec5[5 * at + 3] = px[5 * s + 3];
            rec5[5 * at + 4] = c;
            ts_out[at] = vt[s];
            ++at;
            if (valid_num == (int)((((size_t)(cut_num + 1)) * valid_pcl_size / (size_t)required_cut_num) - 1)) {   /* :79 */
                cut_num++;
                double tmx = ts_out[seg_begin];
                for (long j = seg_begin; j < at; ++j) if (ts_out[j] > tmx) tmx = ts_out[j];
                if (!(tmx < 1.0)) for (long j = seg_begin; j < at; ++j) ts_out[j] = ts_out[j] / tmx;   /* normalize_timestamps :87 */
                if (nseg < max_segments) { seg_sizes[nseg] = at - seg_begin; seg_time[nseg] = last_frame_end_time / (double)1000; }   /* :89 */
                ++nseg;
                last_frame_end_time += (double)c;                                          /* :92 */
                seg_begin = at;
            }
        }
        /* points after the last emitted cut stay in the reference's local split_surface and are dropped; `at` rewinds to them */
        free(idx); free(tmp); free(key);
    }
    free(extracted); free(px); free(vt); free(is_first); free(yaw_fp); free(time_last);
    return status < 0 ? status : (nseg < max_segments ? nseg : max_segments);
}
