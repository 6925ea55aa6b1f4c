// tests/cpp/dropin_main.cpp -- C++ consumer of the drop-in headers (include/limu_dropin), written the way the
// reference's own code uses these classes (L/src/odom_run.cpp:103-106, L/src/tests/hash_map_test.hpp).
// It is compiled against include/limu_dropin FIRST, then the reference's include tree for the shared types
// (common.hpp, utils/types.hpp, lidar/frame.hpp) -- i.e. exactly the include-path swap INTEGRATION.md describes.
// Input/outputs are flat little-endian files so pytest can feed the same data to the oracle:
//   in : int64 n_arrays, then per array: int64 n_doubles, doubles...
//   out: same framing.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "limu/sensors/lidar/icp.hpp"   // resolves to include/limu_dropin/limu/sensors/lidar/icp.hpp (which pulls the drop-in lidar/frame.hpp)

using Arr = std::vector<double>;

static std::vector<Arr> read_all(const char *path) {
    FILE *f = std::fopen(path, "rb");
    if (!f) { std::perror(path); std::exit(2); }
    int64_t n = 0;
    if (std::fread(&n, 8, 1, f) != 1) std::exit(2);
    std::vector<Arr> out(static_cast<size_t>(n));
    for (auto &a : out) {
        int64_t m = 0;
        if (std::fread(&m, 8, 1, f) != 1) std::exit(2);
        a.resize(static_cast<size_t>(m));
        if (m && std::fread(a.data(), 8, static_cast<size_t>(m), f) != static_cast<size_t>(m)) std::exit(2);
    }
    std::fclose(f);
    return out;
}
static void write_all(const char *path, const std::vector<Arr> &arrs) {
    FILE *f = std::fopen(path, "wb");
    int64_t n = static_cast<int64_t>(arrs.size());
    std::fwrite(&n, 8, 1, f);
    for (const auto &a : arrs) {
        int64_t m = static_cast<int64_t>(a.size());
        std::fwrite(&m, 8, 1, f);
        if (m) std::fwrite(a.data(), 8, static_cast<size_t>(m), f);
    }
    std::fclose(f);
}
static utils::Vec3dVector to_points(const Arr &a) {
    utils::Vec3dVector v(a.size() / 3);
    for (size_t i = 0; i < v.size(); ++i) v[i] = utils::Vec3d(a[3 * i], a[3 * i + 1], a[3 * i + 2]);
    return v;
}
static Arr flat(const utils::Vec3dVector &v) {
    Arr a(3 * v.size());
    for (size_t i = 0; i < v.size(); ++i) { a[3 * i] = v[i][0]; a[3 * i + 1] = v[i][1]; a[3 * i + 2] = v[i][2]; }
    return a;
}
static Arr flat(const Sophus::SE3d &T) { return Arr(T.data(), T.data() + 7); }
static Sophus::SE3d to_pose(const Arr &a) { Sophus::SE3d T; for (int i = 0; i < 7; ++i) T.data()[i] = a[static_cast<size_t>(i)]; return T; }

int main(int argc, char **argv) {
    if (argc != 3) { std::fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]); return 2; }
    const auto in = read_all(argv[1]);
    std::vector<Arr> out;
    // in[0] map batch A, in[1] map batch B, in[2] queries, in[3] {tau}, in[4] origin(3), in[5] pose(7) for update
    lidar::VoxelHashMap map(1.0, 30.0, 5);
    out.push_back({map.empty() ? 1.0 : 0.0});
    map.insert_points(to_points(in[0]));
    map.insert_points(to_points(in[1]));
    out.push_back({static_cast<double>(map.size())});
    out.push_back(flat(map.pointcloud()));
    const auto q = to_points(in[2]);
    utils::Vec3dVector nn;
    for (size_t i = 0; i < 64 && i < q.size(); ++i) nn.push_back(map.get_closest_neighbour(q[i]));
    out.push_back(flat(nn));
    const auto corr = map.get_correspondences(q, in[3][0]);
    out.push_back(flat(std::get<0>(corr)));
    out.push_back(flat(std::get<1>(corr)));
    map.remove_points_from_far(utils::Vec3d(in[4][0], in[4][1], in[4][2]));
    out.push_back(flat(map.pointcloud()));
    map.update(to_points(in[1]), to_pose(in[5]));
    out.push_back(flat(map.pointcloud()));
    // in[6] align src, in[7] align tgt, in[8] {th}
    out.push_back(flat(lidar::align_clouds(to_points(in[6]), to_points(in[7]), in[8][0])));
    // in[9] world, in[10] icp source, in[11] init pose, in[12] {tau, kernel, max_iter, eps}
    lidar::VoxelHashMap world(1.0, 100.0, 20);
    world.insert_points(to_points(in[9]));
    out.push_back(flat(lidar::ICP(world, to_points(in[10]), to_pose(in[11]), in[12][0], in[12][1], static_cast<int>(in[12][2]), in[12][3])));
    // KissICP: in[13] = {n_scans}, then per scan: xyz (as doubles holding float values), ts
    auto cfg = std::make_shared<frame::Lidar::ProcessingInfo>();
    cfg->frame_rate = 10.0; cfg->max_range = 100.0; cfg->min_range = 5.0; cfg->min_angle = 0.0; cfg->max_angle = 360.0;
    cfg->num_scan_lines = 16; cfg->frame_split_num = 1; cfg->voxel_size = 1.0; cfg->vox_side_length = 3; cfg->max_points_per_voxel = 10;
    cfg->deskew = true; cfg->min_motion_th = 0.1; cfg->icp_max_iteration = 100; cfg->initial_threshold = 2.0; cfg->estimation_threshold = 0.0001;
    lidar::KissICP kiss(cfg);
    const int n_scans = static_cast<int>(in[13][0]);
    for (int s = 0; s < n_scans; ++s) {
        const Arr &xyz = in[static_cast<size_t>(14 + 2 * s)], &ts = in[static_cast<size_t>(15 + 2 * s)];
        utils::PointCloudXYZI cloud;
        cloud.points.resize(ts.size());
        for (size_t i = 0; i < ts.size(); ++i) {
            cloud.points[i].x = static_cast<float>(xyz[3 * i]); cloud.points[i].y = static_cast<float>(xyz[3 * i + 1]); cloud.points[i].z = static_cast<float>(xyz[3 * i + 2]);
        }
        const auto r = kiss.register_frame(cloud, ts);
        out.push_back(flat(std::get<2>(r)));
        out.push_back({static_cast<double>(std::get<0>(r).size()), static_cast<double>(std::get<1>(r).size()), kiss.has_moved() ? 1.0 : 0.0});
        if (s == n_scans - 1) { out.push_back(flat(std::get<0>(r))); out.push_back(flat(std::get<1>(r))); }
    }
    out.push_back({static_cast<double>(kiss.poses_().size()), static_cast<double>(kiss.local_map_().size())});
    out.push_back(flat(kiss.get_prediction_model()));
    // MotionCompensator on the last scan
    {
        const Arr &xyz = in[static_cast<size_t>(14 + 2 * (n_scans - 1))], &ts = in[static_cast<size_t>(15 + 2 * (n_scans - 1))];
        utils::PointCloudXYZI cloud;
        cloud.points.resize(ts.size());
        for (size_t i = 0; i < ts.size(); ++i) {
            cloud.points[i].x = static_cast<float>(xyz[3 * i]); cloud.points[i].y = static_cast<float>(xyz[3 * i + 1]); cloud.points[i].z = static_cast<float>(xyz[3 * i + 2]);
        }
        lidar::MotionCompensator mc;
        const auto poses = kiss.poses_();
        out.push_back(flat(mc.deskew_scan(cloud, ts, poses[poses.size() - 2], poses[poses.size() - 1])));
    }
    // frame::Lidar (the drop-in lidar/frame.hpp): in[14 + 2 n_scans] = PointCloud2 payload bytes (one per double),
    // in[15 + 2 n_scans] = {message_time, point_step, frame_split_num, messages_to_replay}; the LidarPoint field layout of lidar/frame.hpp:21-23
    {
        const size_t at = static_cast<size_t>(14 + 2 * n_scans);
        if (in.size() > at + 1) {
            const Arr &bytes = in[at], &par = in[at + 1];
            ros::NodeHandle nh;
            frame::Lidar lidar_frame(nh);
            lidar_frame.config->frame_split_num = static_cast<int>(par[2]);
            auto msg = std::make_shared<sensor_msgs::PointCloud2>();
            msg->point_step = static_cast<std::uint32_t>(par[1]);
            msg->height = 1; msg->width = static_cast<std::uint32_t>(bytes.size() / static_cast<size_t>(par[1]));
            msg->data.resize(bytes.size());
            for (size_t i = 0; i < bytes.size(); ++i) msg->data[i] = static_cast<std::uint8_t>(bytes[i]);
            const char *names[6] = {"x", "y", "z", "intensity", "ring", "timestamp"};
            const int offs[6] = {0, 4, 8, 12, 14, 16}, types[6] = {7, 7, 7, 2, 4, 8};
            for (int k = 0; k < 6; ++k) {
                sensor_msgs::PointField f;
                f.name = names[k]; f.offset = static_cast<std::uint32_t>(offs[k]); f.datatype = static_cast<std::uint8_t>(types[k]); f.count = 1;
                msg->fields.push_back(f);
            }
            const int replay = static_cast<int>(par[3]);   // the same payload `replay` times: the 20th message onwards is split (frame.cpp:64)
            Arr sizes, times, rec, tss;
            for (int r = 0; r < replay; ++r) {
                msg->header.stamp.fromSec(par[0]);
                lidar_frame.initialize(msg);
                lidar_frame.process_frame();
                while (!lidar_frame.buffer_empty()) {
                    const auto cloud = lidar_frame.get_lidar_buffer_front();
                    const auto t = lidar_frame.get_segment_ts_front();
                    if (r == replay - 1) {
                        sizes.push_back(static_cast<double>(cloud->points.size()));
                        times.push_back(lidar_frame.curr_acc_segment_time());
                        for (size_t i = 0; i < cloud->points.size(); ++i) {
                            const auto &q = cloud->points[i];
                            rec.insert(rec.end(), {q.x, q.y, q.z, q.intensity, q.curvature});
                            tss.push_back(t[i]);
                        }
                    }
                    lidar_frame.pop();
                }
            }
            out.push_back(sizes); out.push_back(times); out.push_back(rec); out.push_back(tss);
        }
    }
    write_all(argv[2], out);
    std::printf("dropin_test ok: %zu output arrays\n", out.size());
    return 0;
}
