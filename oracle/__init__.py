"""ORACLE bindings -- test infrastructure, not product code.

Two CPU implementations of the reference hot path behind ONE numpy-facing API, so a test can run the
same check against either:

* ``load_ref(mt=False)``  -> oracle/_ref/liblimu_ref(.so|_mt.so): the reference's UNMODIFIED sources
  (Oreoluwa-Se/Lidar-Imu-Slam env_ws/src/limu) compiled by oracle/Makefile against the vendored
  Eigen/Sophus tarballs and the shims in oracle/shims. Prebuilt in the authoring container; travels to
  the GPU box as a binary (``/root/reference`` does not exist there).
* ``load_port()``        -> oracle/build/liblimu_oracle.so: the plain-C restatement (limu_oracle.c),
  buildable anywhere gcc exists.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's ``cpu_baseline`` / ``--impl reference`` legs may
import this package. The product (``lidar-imu-slam_b200``) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "liblimu_ref.so")
REF_MT_SO = os.path.join(HERE, "_ref", "liblimu_ref_mt.so")
PORT_SO = os.path.join(HERE, "build", "liblimu_oracle.so")

_dp = C.POINTER(C.c_double)
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)
_lp = C.POINTER(C.c_long)


def _d(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _f(a):
    return a.ctypes.data_as(_fp) if a is not None else None


def _i(a):
    return a.ctypes.data_as(_ip) if a is not None else None


def _l(a):
    return a.ctypes.data_as(_lp) if a is not None else None


def _pts(x):
    x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, 3)
    return x


def build_port(force: bool = False) -> str:
    """Compile the C restatement (gcc only)."""
    srcs = [os.path.join(HERE, f) for f in ("limu_oracle.c", "limu_oracle_frame.c", "limu_oracle.h")]
    if force or not os.path.exists(PORT_SO) or os.path.getmtime(PORT_SO) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-s", "-C", HERE, "port"])
    return PORT_SO


def build_ref() -> bool:
    """Compile the reference itself; only possible where /root/reference exists."""
    if not os.path.isdir("/root/reference/env_ws/src/limu"):
        return os.path.exists(REF_SO)
    subprocess.check_call(["make", "-s", "-j2", "-C", HERE, "ref"])
    return True


def have_ref() -> bool:
    return os.path.exists(REF_SO)


class _Api:
    """Uniform API. ``p`` is the symbol prefix ('ref_' or 'lo_'); ``kind`` is 'reference' or 'port'."""

    def __init__(self, lib, p, kind):
        self.lib, self.p, self.kind = lib, p, kind
        f = self._f
        f("vox_index", None, [_dp, C.c_long, C.c_double, _ip])
        f("transform", None, [_dp, _dp, C.c_long])
        f("se3_exp", None, [_dp, _dp])
        f("se3_log", None, [_dp, _dp])
        f("se3_mul", None, [_dp, _dp, _dp])
        f("se3_inv", None, [_dp, _dp])
        f("delta_pose", None, [_dp, _dp, _dp])
        f("map_create", C.c_void_p, [C.c_double, C.c_double, C.c_int])
        f("map_destroy", None, [C.c_void_p])
        f("map_clear", None, [C.c_void_p])
        f("map_empty", C.c_int, [C.c_void_p])
        f("map_num_voxels", C.c_long, [C.c_void_p])
        f("map_insert", None, [C.c_void_p, _dp, C.c_long])
        f("map_update", None, [C.c_void_p, _dp, C.c_long, _dp])
        f("map_remove_far", None, [C.c_void_p, _dp])
        f("map_dump", C.c_long, [C.c_void_p, _ip, _ip, _dp, C.c_long, C.c_long, _lp])
        f("deskew", None, [_fp, _dp, C.c_long, _dp, _dp, _dp])
        if kind == "port":
            f("deskew_imu", None, [_fp, _fp, C.c_long, _dp, C.c_long, _dp, _dp, _dp, _dp])
        elif hasattr(lib, "ref_imu_deskew"):
            f("imu_deskew", C.c_long, [_fp, _fp, C.c_long, _dp, C.c_long, C.c_double, _dp, _dp, _dp, _dp, _dp, C.c_long, _dp, _dp, _fp])
            if hasattr(lib, "ref_imu_set_last_lidar_end_time"):
                f("imu_set_last_lidar_end_time", None, [C.c_double])
        if kind == "reference":
            f("map_closest", None, [C.c_void_p, _dp, C.c_long, _dp])
            f("map_correspondences", C.c_long, [C.c_void_p, _dp, C.c_long, C.c_double, _dp, _dp])
            f("align", None, [_dp, _dp, C.c_long, C.c_double, _dp])
            f("icp", None, [C.c_void_p, _dp, C.c_long, _dp, C.c_double, C.c_double, C.c_int, C.c_double, _dp])
            f("icp_trace", C.c_int, [C.c_void_p, _dp, C.c_long, _dp, C.c_double, C.c_double, C.c_int, C.c_double, _dp, _dp, _lp, _dp])
            f("threshold_create", C.c_void_p, [C.c_double, C.c_double, C.c_double])
            f("threshold_destroy", None, [C.c_void_p])
            f("threshold_step", C.c_double, [C.c_void_p, _dp])
            f("voxel_downsample", C.c_long, [C.c_void_p, _dp, C.c_long, C.c_double, _dp])
            f("voxelize", None, [C.c_void_p, _dp, C.c_long, C.c_double, _dp, _lp, _dp, _lp])
            f("iqr", C.c_long, [C.c_void_p, _dp, C.c_long, _dp])
            f("num_threads", C.c_int, [])
        else:
            f("map_closest", None, [C.c_void_p, _dp, C.c_long, _dp, _ip, _ip])
            f("map_correspondences", C.c_long, [C.c_void_p, _dp, C.c_long, C.c_double, _dp, _dp, _lp])
            f("align", None, [_dp, _dp, C.c_long, C.c_double, _dp, _dp, _dp, _dp])
            f("icp", C.c_int, [C.c_void_p, _dp, C.c_long, _dp, C.c_double, C.c_double, C.c_int, C.c_double, _dp, _dp, _lp, _dp, _dp])
            f("threshold_init", None, [C.c_void_p, C.c_double, C.c_double, C.c_double])
            f("threshold_step", C.c_double, [C.c_void_p, _dp])
            f("voxel_downsample", C.c_long, [_dp, C.c_long, C.c_double, _dp, _lp])
            f("voxelize", None, [_dp, C.c_long, C.c_double, _dp, _lp, _dp, _lp])
            f("iqr", C.c_long, [_dp, C.c_long, _dp, _dp])
            f("kiss_last_iterations", C.c_int, [C.c_void_p])
            f("kiss_last_sigma", C.c_double, [C.c_void_p])
            f("kiss_map", C.c_void_p, [C.c_void_p])
            f("map_set_mode", None, [C.c_void_p, C.c_int])
            f("plane_normal", C.c_int, [_dp, C.c_int, _dp])
            f("kiss_set_mode", None, [C.c_void_p, C.c_int])
        if hasattr(lib, p + "process_frame"):
            f("process_frame", C.c_long, [C.c_char_p, C.c_long, C.c_int, C.c_int, C.c_char_p, _ip, _ip, _ip, _dp, C.c_double, C.c_int, C.c_long, _lp, _dp, _fp, _dp])
        f("kiss_create", C.c_void_p, [C.c_double, C.c_double, C.c_int, C.c_int, C.c_double, C.c_int, C.c_double, C.c_double])
        f("kiss_destroy", None, [C.c_void_p])
        f("kiss_register_points", None, [C.c_void_p, _dp, C.c_long, _dp, _lp, _dp, _lp, _dp])
        f("kiss_register_cloud", None, [C.c_void_p, _fp, _dp, C.c_long, _dp, _lp, _dp, _lp, _dp])
        f("kiss_num_poses", C.c_long, [C.c_void_p])
        f("kiss_pose", None, [C.c_void_p, C.c_long, _dp])
        self._scratch_kiss = None

    def _f(self, name, res, args):
        fn = getattr(self.lib, self.p + name)
        fn.restype, fn.argtypes = res, args
        setattr(self, "_" + name, fn)

    # -- utils ---------------------------------------------------------------------------------------
    def num_threads(self) -> int:
        return int(self._num_threads()) if self.kind == "reference" else 1

    def vox_index(self, xyz, v):
        xyz = _pts(xyz)
        keys = np.empty((len(xyz), 3), np.int32)
        self._vox_index(_d(xyz), len(xyz), float(v), _i(keys))
        return keys

    def transform(self, pose7, xyz):
        out = _pts(xyz).copy()
        pose7 = np.ascontiguousarray(pose7, np.float64)
        if len(out):
            self._transform(_d(pose7), _d(out), len(out))
        return out

    def se3_exp(self, x6):
        x6 = np.ascontiguousarray(x6, np.float64)
        out = np.empty(7)
        self._se3_exp(_d(x6), _d(out))
        return out

    def se3_log(self, p7):
        p7 = np.ascontiguousarray(p7, np.float64)
        out = np.empty(6)
        self._se3_log(_d(p7), _d(out))
        return out

    def se3_mul(self, a, b):
        a = np.ascontiguousarray(a, np.float64)
        b = np.ascontiguousarray(b, np.float64)
        out = np.empty(7)
        self._se3_mul(_d(a), _d(b), _d(out))
        return out

    def se3_inv(self, a):
        a = np.ascontiguousarray(a, np.float64)
        out = np.empty(7)
        self._se3_inv(_d(a), _d(out))
        return out

    def delta_pose(self, a, b):
        a = np.ascontiguousarray(a, np.float64)
        b = np.ascontiguousarray(b, np.float64)
        out = np.empty(6)
        self._delta_pose(_d(a), _d(b), _d(out))
        return out

    # -- free functions ----------------------------------------------------------------------------------
    def deskew(self, xyz_f32, ts, T0, T1):
        x = np.ascontiguousarray(xyz_f32, np.float32).reshape(-1, 3)
        ts = np.ascontiguousarray(ts, np.float64)
        T0 = np.ascontiguousarray(T0, np.float64)
        T1 = np.ascontiguousarray(T1, np.float64)
        out = np.empty((len(x), 3))
        self._deskew(_f(x), _d(ts), len(x), _d(T0), _d(T1), _d(out))
        return out

    def imu_set_last_lidar_end_time(self, t):
        """REFERENCE ONLY: EKF::last_lidar_end_time as the next imu_deskew_reference call finds it (0 for a fresh filter)."""
        self._imu_set_last_lidar_end_time(float(t))

    def imu_deskew_reference(self, xyz_f32, curv_ms, imu, lidar_beg_time, mean_acc, p_imu_lidar, gyro_bias):
        """REFERENCE ONLY: run kalman::EKF::motion_compensation_with_imu (ekf.cpp:292-469) on one scan + IMU window.
        Returns dict(deskewed [n,3] f64, written_back [n,3] f32, table [M,22], rot_end [9], pos_lidar_end [3])."""
        assert self.kind == "reference"
        x = np.ascontiguousarray(xyz_f32, np.float32).reshape(-1, 3)
        c = np.ascontiguousarray(curv_ms, np.float32)
        imu = np.ascontiguousarray(imu, np.float64).reshape(-1, 7)
        out = np.empty((len(x), 3))
        wb = np.empty((len(x), 3), np.float32)
        table = np.zeros((len(imu) + 4, 22))
        rot_end, ple = np.empty(9), np.empty(3)
        ma, pil, gb = (np.ascontiguousarray(v, np.float64) for v in (mean_acc, p_imu_lidar, gyro_bias))
        M = self._imu_deskew(_f(x), _f(c), len(x), _d(imu), len(imu), float(lidar_beg_time), _d(ma), _d(pil), _d(gb), _d(out), _d(table), len(table),
                             _d(rot_end), _d(ple), _f(wb))
        return {"deskewed": out, "written_back": wb, "table": table[:M].copy(), "rot_end": rot_end, "pos_lidar_end": ple}

    def deskew_imu(self, xyz_f32, curv_ms, table, rot_end, pos_lidar_end, p_imu_lidar):
        """PORT ONLY: per-point loop of motion_compensation_with_imu (ekf.cpp:420-468). Returns (deskewed f64, written-back f32)."""
        assert self.kind == "port"
        x = np.ascontiguousarray(xyz_f32, np.float32).reshape(-1, 3).copy()
        c = np.ascontiguousarray(curv_ms, np.float32)
        t = np.ascontiguousarray(table, np.float64).reshape(-1, 22)
        out = np.empty((len(x), 3))
        re, pe, pil = (np.ascontiguousarray(v, np.float64) for v in (rot_end, pos_lidar_end, p_imu_lidar))
        self._deskew_imu(_f(x), _f(c), len(x), _d(t), len(t), _d(re), _d(pe), _d(pil), _d(out))
        return out, x

    def process_frame(self, data, fields, cfg, message_time, scan_count, max_segments=16):
        """frame::Lidar::process_frame (lidar/frame.cpp:101-193) on one PointCloud2 payload.
        data: uint8 [n, point_step]; fields: [(name, offset, PointField datatype, count)];
        cfg: dict(min_range, max_range, min_angle, max_angle, frame_rate, num_scan_lines, frame_split_num).
        Returns a list of segments dict(points [m,5] f32 = x,y,z,intensity,curvature; ts [m] f64; time) or a negative status."""
        data = np.ascontiguousarray(data, np.uint8)
        n, step = data.shape
        names = b"".join(f[0].encode() + b"\0" for f in fields)
        offs = np.array([f[1] for f in fields], np.int32)
        dts = np.array([f[2] for f in fields], np.int32)
        cnts = np.array([f[3] for f in fields], np.int32)
        c = np.array([cfg["min_range"], cfg["max_range"], cfg["min_angle"], cfg["max_angle"], cfg["frame_rate"], cfg["num_scan_lines"],
                      cfg["frame_split_num"]], np.float64)
        seg_sizes = np.zeros(max_segments, np.int64)
        seg_time = np.zeros(max_segments)
        rec5 = np.zeros((max(n, 1), 5), np.float32)
        ts = np.zeros(max(n, 1))
        k = self._process_frame(data.ctypes.data_as(C.c_char_p), n, step, len(fields), names, _i(offs), _i(dts), _i(cnts), _d(c), float(message_time),
                                int(scan_count), max_segments, _l(seg_sizes), _d(seg_time), _f(rec5), _d(ts))
        if k < 0:
            return int(k)
        out, at = [], 0
        for j in range(k):
            m = int(seg_sizes[j])
            out.append({"points": rec5[at:at + m].copy(), "ts": ts[at:at + m].copy(), "time": float(seg_time[j])})
            at += m
        return out

    def plane_normal(self, pts):
        """PORT ONLY: the voxel plane of the opt-in point-to-plane variant. Returns the unit normal or None when not planar."""
        pts = _pts(pts)
        n = np.zeros(3)
        return n if self._plane_normal(_d(pts), len(pts), _d(n)) else None

    def align(self, src, tgt, th):
        """Returns dict(pose=7) for the reference; the port adds H (6x6), g (6), x (6)."""
        src, tgt = _pts(src), _pts(tgt)
        pose = np.empty(7)
        if self.kind == "reference":
            self._align(_d(src), _d(tgt), len(src), float(th), _d(pose))
            return {"pose": pose}
        H, g, x = np.empty((6, 6)), np.empty(6), np.empty(6)
        self._align(_d(src), _d(tgt), len(src), float(th), _d(H), _d(g), _d(x), _d(pose))
        return {"pose": pose, "H": H, "g": g, "x": x}

    def _kiss_scratch(self):
        if self._scratch_kiss is None:
            self._scratch_kiss = self._kiss_create(1.0, 100.0, 10, 0, 0.1, 500, 2.0, 1e-4)
        return self._scratch_kiss

    def voxel_downsample(self, xyz, s):
        xyz = _pts(xyz)
        out = np.empty((max(len(xyz), 1), 3))
        if self.kind == "reference":
            n = self._voxel_downsample(self._kiss_scratch(), _d(xyz), len(xyz), float(s), _d(out))
        else:
            n = self._voxel_downsample(_d(xyz), len(xyz), float(s), _d(out), None)
        return out[:n].copy()

    def iqr(self, xyz):
        xyz = _pts(xyz)
        out = np.empty((max(len(xyz), 1), 3))
        if self.kind == "reference":
            n = self._iqr(self._kiss_scratch(), _d(xyz), len(xyz), _d(out))
        else:
            n = self._iqr(_d(xyz), len(xyz), _d(out), None)
        return out[:n].copy()

    def voxelize(self, xyz, v):
        xyz = _pts(xyz)
        src = np.empty((max(len(xyz), 1), 3))
        down = np.empty((max(len(xyz), 1), 3))
        ns, nd = C.c_long(0), C.c_long(0)
        if self.kind == "reference":
            self._voxelize(self._kiss_scratch(), _d(xyz), len(xyz), float(v), _d(src), C.byref(ns), _d(down), C.byref(nd))
        else:
            self._voxelize(_d(xyz), len(xyz), float(v), _d(src), C.byref(ns), _d(down), C.byref(nd))
        return src[: ns.value].copy(), down[: nd.value].copy()

    def Map(self, vox_size, max_distance, cap):
        return _Map(self, vox_size, max_distance, cap)

    def Threshold(self, init_th, min_motion, max_range):
        return _Threshold(self, init_th, min_motion, max_range)

    def Kiss(self, **cfg):
        return _Kiss(self, **cfg)

    def icp(self, m, xyz, init7, tau, th, max_iter, eps, trace=False):
        """Returns dict(pose, iters[, est (iters x 7), ncorr (iters), hg (iters x 42, port only), src_after])."""
        xyz = _pts(xyz)
        init7 = np.ascontiguousarray(init7, np.float64)
        pose = np.empty(7)
        if not trace and self.kind == "reference":
            self._icp(m.h, _d(xyz), len(xyz), _d(init7), tau, th, max_iter, eps, _d(pose))
            return {"pose": pose}
        est = np.zeros((max_iter, 7))
        nc = np.zeros(max_iter, dtype=np.int64)
        after = np.empty((max(len(xyz), 1), 3))
        if self.kind == "reference":
            it = self._icp_trace(m.h, _d(xyz), len(xyz), _d(init7), tau, th, max_iter, eps, _d(pose), _d(est), _l(nc), _d(after))
            return {"pose": pose, "iters": it, "est": est[:it], "ncorr": nc[:it], "src_after": after[: len(xyz)]}
        hg = np.zeros((max_iter, 42))
        it = self._icp(m.h, _d(xyz), len(xyz), _d(init7), tau, th, max_iter, eps, _d(pose), _d(est), _l(nc), _d(hg), _d(after))
        return {"pose": pose, "iters": it, "est": est[:it], "ncorr": nc[:it], "hg": hg[:it], "src_after": after[: len(xyz)]}


class _Map:
    def __init__(self, api, vox_size, max_distance, cap, handle=None):
        self.api, self.cap = api, cap
        self.owned = handle is None
        self.h = api._map_create(float(vox_size), float(max_distance), int(cap)) if handle is None else handle

    def __del__(self):
        if getattr(self, "owned", False) and self.h:
            self.api._map_destroy(self.h)
            self.h = None

    def set_mode(self, icp_mode):
        """PORT ONLY: opt-in neighbour rule (1 = nearest of the 27-cell neighbourhood)."""
        self.api._map_set_mode(self.h, int(icp_mode))

    def insert(self, xyz):
        xyz = _pts(xyz)
        self.api._map_insert(self.h, _d(xyz), len(xyz))

    def update(self, xyz, pose7):
        xyz = _pts(xyz)
        pose7 = np.ascontiguousarray(pose7, np.float64)
        self.api._map_update(self.h, _d(xyz), len(xyz), _d(pose7))

    def remove_far(self, origin):
        o = np.ascontiguousarray(origin, np.float64)
        self.api._map_remove_far(self.h, _d(o))

    def clear(self):
        self.api._map_clear(self.h)

    def empty(self):
        return bool(self.api._map_empty(self.h))

    def num_voxels(self):
        return int(self.api._map_num_voxels(self.h))

    def closest(self, xyz, with_index=False):
        xyz = _pts(xyz)
        out = np.empty((len(xyz), 3))
        if self.api.kind == "reference":
            self.api._map_closest(self.h, _d(xyz), len(xyz), _d(out))
            return out
        key = np.empty((len(xyz), 3), np.int32)
        rank = np.empty(len(xyz), np.int32)
        self.api._map_closest(self.h, _d(xyz), len(xyz), _d(out), _i(key), _i(rank))
        return (out, key, rank) if with_index else out

    def correspondences(self, xyz, tau, with_index=False):
        xyz = _pts(xyz)
        src = np.empty((max(len(xyz), 1), 3))
        tgt = np.empty((max(len(xyz), 1), 3))
        if self.api.kind == "reference":
            n = self.api._map_correspondences(self.h, _d(xyz), len(xyz), float(tau), _d(src), _d(tgt))
            return src[:n].copy(), tgt[:n].copy()
        idx = np.empty(max(len(xyz), 1), np.int64)
        n = self.api._map_correspondences(self.h, _d(xyz), len(xyz), float(tau), _d(src), _d(tgt), _l(idx))
        return (src[:n].copy(), tgt[:n].copy(), idx[:n].copy()) if with_index else (src[:n].copy(), tgt[:n].copy())

    def dump(self):
        """(keys [V,3] int32, counts [V] int32, pts [sum(counts),3]) in creation order."""
        npts = C.c_long(0)
        nv = self.api._map_dump(self.h, None, None, None, 0, 0, C.byref(npts))
        keys = np.empty((max(nv, 1), 3), np.int32)
        counts = np.empty(max(nv, 1), np.int32)
        pts = np.empty((max(npts.value, 1), 3))
        self.api._map_dump(self.h, _i(keys), _i(counts), _d(pts), nv, npts.value, C.byref(npts))
        return keys[:nv], counts[:nv], pts[: npts.value]


class _Threshold:
    def __init__(self, api, init_th, min_motion, max_range):
        self.api = api
        if api.kind == "reference":
            self.h = api._threshold_create(init_th, min_motion, max_range)
        else:
            self.buf = (C.c_double * 16)()
            self.h = C.cast(self.buf, C.c_void_p)
            api._threshold_init(self.h, init_th, min_motion, max_range)

    def step(self, dev7):
        dev7 = np.ascontiguousarray(dev7, np.float64)
        return float(self.api._threshold_step(self.h, _d(dev7)))

    def __del__(self):
        if self.api.kind == "reference" and getattr(self, "h", None):
            self.api._threshold_destroy(self.h)
            self.h = None


class _Kiss:
    """KissICP pipeline object (icp.hpp:31-68)."""

    def __init__(self, api, voxel_size=1.0, max_range=100.0, cap=10, deskew=False, min_motion_th=0.1,
                 icp_max_iteration=500, initial_threshold=2.0, estimation_threshold=1e-4):
        self.api = api
        self.h = api._kiss_create(voxel_size, max_range, cap, int(deskew), min_motion_th, icp_max_iteration,
                                  initial_threshold, estimation_threshold)

    def __del__(self):
        if getattr(self, "h", None):
            self.api._kiss_destroy(self.h)
            self.h = None

    def set_mode(self, icp_mode):
        """PORT ONLY: opt-in neighbour rule of the local map."""
        self.api._kiss_set_mode(self.h, int(icp_mode))

    def _out(self, n):
        return np.empty((max(n, 1), 3)), np.empty((max(n, 1), 3)), C.c_long(0), C.c_long(0), np.empty(7)

    def register_points(self, xyz):
        xyz = _pts(xyz)
        down, src, nd, ns, pose = self._out(len(xyz))
        self.api._kiss_register_points(self.h, _d(xyz), len(xyz), _d(down), C.byref(nd), _d(src), C.byref(ns), _d(pose))
        return down[: nd.value].copy(), src[: ns.value].copy(), pose

    def register_cloud(self, xyz_f32, ts):
        x = np.ascontiguousarray(xyz_f32, np.float32).reshape(-1, 3)
        ts = np.ascontiguousarray(ts, np.float64)
        down, src, nd, ns, pose = self._out(len(x))
        self.api._kiss_register_cloud(self.h, _f(x), _d(ts), len(x), _d(down), C.byref(nd), _d(src), C.byref(ns), _d(pose))
        return down[: nd.value].copy(), src[: ns.value].copy(), pose

    def poses(self):
        n = self.api._kiss_num_poses(self.h)
        out = np.empty((n, 7))
        for i in range(n):
            self.api._kiss_pose(self.h, i, _d(out[i]))
        return out

    def last_iterations(self):
        return int(self.api._kiss_last_iterations(self.h)) if self.api.kind == "port" else -1

    def last_sigma(self):
        return float(self.api._kiss_last_sigma(self.h)) if self.api.kind == "port" else float("nan")

    def map(self):
        if self.api.kind != "port":
            raise NotImplementedError
        m = _Map(self.api, 0, 0, 0, handle=self.api._kiss_map(self.h))
        return m


_cache = {}


def load_ref(mt: bool = False) -> _Api:
    key = "ref_mt" if mt else "ref"
    if key not in _cache:
        path = REF_MT_SO if mt else REF_SO
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: build it where /root/reference exists (make -C oracle ref)")
        _cache[key] = _Api(C.CDLL(path), "ref_", "reference")
    return _cache[key]


def load_port() -> _Api:
    if "port" not in _cache:
        build_port()
        _cache["port"] = _Api(C.CDLL(PORT_SO), "lo_", "port")
    return _cache["port"]


class RefEkf:
    """The reference's own kalman::EKF (compiled, oracle/ref_ekf_driver.cpp) behind the same calls as limu_b200.Ekf."""

    ORDER = ("lidar_pose_trail", "noise_scale", "init_pos_noise", "init_vel_noise", "init_ori_noise", "init_bga_noise", "init_baa_noise", "init_bat_noise",
             "acc_process_noise", "gyro_process_noise", "acc_process_noise_rev", "gyro_process_noise_rev", "init_lidar_imu_time_noise",
             "init_pos_trail_noise", "init_ori_trail_noise", "visualZuptR")
    DEFAULTS = dict(lidar_pose_trail=20, noise_scale=1.0, init_pos_noise=1e-3, init_vel_noise=1e-3, init_ori_noise=1e-3, init_bga_noise=1e-3, init_baa_noise=1e-3,
                    init_bat_noise=1e-3, acc_process_noise=0.03, gyro_process_noise=0.00017, acc_process_noise_rev=0.03, gyro_process_noise_rev=0.00017,
                    init_lidar_imu_time_noise=1e-3, init_pos_trail_noise=1e-3, init_ori_trail_noise=1e-3, visualZuptR=1e-3)

    def __init__(self, **params):
        self.lib = load_ref().lib
        L = self.lib
        L.ref_ekf_create.restype = C.c_void_p
        L.ref_ekf_create.argtypes = [C.POINTER(C.c_double)]
        L.ref_ekf_dim.restype = C.c_long
        for name, args in (("ref_ekf_destroy", [C.c_void_p]), ("ref_ekf_dim", [C.c_void_p]), ("ref_ekf_get", [C.c_void_p] + [C.POINTER(C.c_double)] * 3),
                           ("ref_ekf_set", [C.c_void_p] + [C.POINTER(C.c_double)] * 2), ("ref_ekf_predict", [C.c_void_p, C.c_double] + [C.POINTER(C.c_double)] * 5),
                           ("ref_ekf_normalize", [C.c_void_p, C.c_int]), ("ref_ekf_zero_vel_update", [C.c_void_p, C.c_double]), ("ref_ekf_augment", [C.c_void_p]),
                           ("ref_ekf_undo_augmentation", [C.c_void_p])):
            getattr(L, name).argtypes = args
        p = dict(self.DEFAULTS)
        p.update(params)
        arr = np.array([float(p[k]) for k in self.ORDER])
        self.h = L.ref_ekf_create(_d(arr))
        self.dim = int(L.ref_ekf_dim(self.h))

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.ref_ekf_destroy(self.h)
            self.h = None

    def state(self):
        m, P, t = np.empty(self.dim), np.empty((self.dim, self.dim)), np.empty(1)
        self.lib.ref_ekf_get(self.h, _d(m), _d(P), _d(t))
        return m, P, float(t[0])

    def set_state(self, m=None, P=None):
        self.lib.ref_ekf_set(self.h, _d(np.ascontiguousarray(m, np.float64)) if m is not None else None, _d(np.ascontiguousarray(P, np.float64)) if P is not None else None)

    def predict(self, t, xg, xa, calc_grav, trans_lidar_imu, rot_lidar_imu):
        a = [np.ascontiguousarray(x, np.float64) for x in (xg, xa, calc_grav, trans_lidar_imu, np.asarray(rot_lidar_imu, np.float64).reshape(9))]
        self.lib.ref_ekf_predict(self.h, float(t), *[_d(x) for x in a])

    def normalize_quaternions(self, only_current=False):
        self.lib.ref_ekf_normalize(self.h, int(only_current))

    def zero_velocity_update(self, r):
        self.lib.ref_ekf_zero_vel_update(self.h, float(r))

    def augment_pose_trail(self):
        self.lib.ref_ekf_augment(self.h)

    def undo_augmentation(self):
        self.lib.ref_ekf_undo_augmentation(self.h)
