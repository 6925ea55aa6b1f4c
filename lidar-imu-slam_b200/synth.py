"""Synthetic spinning-LiDAR scans for tests and bench.py (the reference ships no
input data -- its rosbag is git-ignored -- so the workload is generated, seeded and deterministic).

Scene: ground plane + axis-aligned boxes + vertical cylinders (non-symmetric, seeded). A scan is ray-cast
from a sensor that MOVES during the sweep (pose interpolated between the scan's start and end pose), so
constant-velocity deskewing has something to undo. Output rows are float32 (x, y, z, t) in the sensor
frame at each ray's firing time, t in [0,1) = azimuth fraction -- the layout liblimu_cuda takes.
"""
from __future__ import annotations

import numpy as np


class Scene:
    def __init__(self, seed=42, n_boxes=28, n_cyl=14, extent=90.0, street=False):
        rng = np.random.default_rng(seed)
        if street:   # urban canyon: two facade rows along x, clutter between them
            ys = np.concatenate([np.full(n_boxes // 2, 14.0), np.full(n_boxes - n_boxes // 2, -15.0)]) + rng.normal(size=n_boxes) * 1.5
            xs = rng.uniform(-extent, extent, n_boxes)
            c = np.stack([xs, ys], 1)
            half = np.stack([rng.uniform(4, 12, n_boxes), rng.uniform(2, 5, n_boxes)], 1)
        else:
            c = rng.uniform(-extent, extent, (n_boxes, 2))
            half = rng.uniform(1.0, 7.0, (n_boxes, 2))
        h = rng.uniform(2.0, 18.0, n_boxes)
        self.box_lo = np.concatenate([c - half, np.zeros((n_boxes, 1))], 1)
        self.box_hi = np.concatenate([c + half, h[:, None]], 1)
        self.cyl_c = rng.uniform(-extent, extent, (n_cyl, 2))
        self.cyl_r = rng.uniform(0.2, 1.2, n_cyl)
        self.cyl_h = rng.uniform(3.0, 12.0, n_cyl)
        self.sensor_height = 1.8


def _rot_z(yaw):
    c, s = np.cos(yaw), np.sin(yaw)
    z, o = np.zeros_like(c), np.ones_like(c)
    return np.stack([np.stack([c, -s, z], -1), np.stack([s, c, z], -1), np.stack([z, z, o], -1)], -2)


def loop_trajectory(n, radius=30.0, step=1.0):
    """n poses (x, y, yaw) on a circle of `radius`, `step` metres of arc apart, yaw along the tangent."""
    a = np.arange(n) * (step / radius)
    return np.stack([radius * np.sin(a), radius * (1 - np.cos(a)), a], 1)


def pose7_of(xyyaw, height=1.8):
    """(x, y, yaw) -> pose {qx,qy,qz,qw,tx,ty,tz}."""
    x, y, yaw = xyyaw
    return np.array([0.0, 0.0, np.sin(yaw / 2), np.cos(yaw / 2), x, y, height])


def cast_scan(scene, pose_start, pose_end, beams=64, azimuth_steps=2000, elev=(-25.0, 2.0), rng_gate=(5.0, 100.0),
              noise=0.02, seed=0, device="cpu"):
    """Ray-cast one sweep. pose_* = (x, y, yaw) world poses at sweep start / end. The arithmetic runs in
    torch float64 on `device` (bench.py uses the GPU to build its workload quickly; the range noise is
    always drawn on the host from `seed`, so a scan depends on the device only through libm rounding)."""
    import torch
    dev = torch.device(device)
    f64 = dict(dtype=torch.float64, device=dev)
    el = torch.deg2rad(torch.linspace(elev[0], elev[1], beams, **f64))
    az_i = torch.arange(azimuth_steps, **f64)
    t = (az_i / azimuth_steps).repeat_interleave(beams)            # time-ordered: all beams fire per azimuth step
    az = (az_i * (2 * np.pi / azimuth_steps)).repeat_interleave(beams)
    e = el.repeat(azimuth_steps)
    d_s = torch.stack([torch.cos(e) * torch.cos(az), torch.cos(e) * torch.sin(az), torch.sin(e)], 1)   # sensor-frame directions
    p0 = torch.as_tensor(np.asarray(pose_start, float), **f64)
    p1 = torch.as_tensor(np.asarray(pose_end, float), **f64)
    xy = p0[None, :2] + t[:, None] * (p1[:2] - p0[:2])[None]
    yaw = p0[2] + t * (p1[2] - p0[2])
    c, s = torch.cos(yaw), torch.sin(yaw)
    d_w = torch.stack([c * d_s[:, 0] - s * d_s[:, 1], s * d_s[:, 0] + c * d_s[:, 1], d_s[:, 2]], 1)
    o_w = torch.cat([xy, torch.full((len(t), 1), scene.sensor_height, **f64)], 1)
    inf = torch.tensor(float("inf"), **f64)
    # ground z = 0
    tg = -o_w[:, 2] / d_w[:, 2]
    best = torch.where(tg > 0, tg, inf)
    # boxes (slab test), chunked to bound memory
    inv = 1.0 / d_w
    lo_all = torch.as_tensor(scene.box_lo, **f64)
    hi_all = torch.as_tensor(scene.box_hi, **f64)
    for b0 in range(0, len(lo_all), 8):
        lo, hi = lo_all[b0:b0 + 8], hi_all[b0:b0 + 8]
        t1 = (lo[None] - o_w[:, None]) * inv[:, None]
        t2 = (hi[None] - o_w[:, None]) * inv[:, None]
        tn = torch.minimum(t1, t2).nan_to_num(nan=-float("inf")).amax(dim=2)
        tf = torch.maximum(t1, t2).nan_to_num(nan=float("inf")).amin(dim=2)
        hit = (tf >= tn) & (tf > 0) & (tn > 0)
        best = torch.minimum(best, torch.where(hit, tn, inf).amin(dim=1))
    # vertical cylinders
    for cc, r, h in zip(scene.cyl_c, scene.cyl_r, scene.cyl_h):
        oc = o_w[:, :2] - torch.as_tensor(cc, **f64)
        a = (d_w[:, :2] ** 2).sum(1)
        b = 2 * (oc * d_w[:, :2]).sum(1)
        c0 = (oc ** 2).sum(1) - r * r
        disc = b * b - 4 * a * c0
        tc = (-b - torch.sqrt(disc.clamp_min(0))) / (2 * a)
        z = o_w[:, 2] + tc * d_w[:, 2]
        ok = (disc > 0) & (tc > 0) & (z >= 0) & (z <= h)
        best = torch.minimum(best, torch.where(ok, tc, inf))
    nz = torch.as_tensor(np.random.default_rng(seed).normal(size=len(t)) * noise, **f64)
    rngs = best + nz
    valid = torch.isfinite(best) & (rngs >= rng_gate[0]) & (rngs <= rng_gate[1])   # range gate lidar/frame.hpp:65-66
    rngs = torch.where(valid, rngs, torch.zeros_like(rngs))
    out = torch.cat([d_s * rngs[:, None], t[:, None]], 1)[valid].to(torch.float32)
    return np.ascontiguousarray(out.cpu().numpy())


def pad_scan(scan, n, seed=0):
    """Resize a scan to exactly n points (subsample, or repeat with sub-centimetre jitter), keeping time order."""
    if len(scan) == n:
        return scan
    rng = np.random.default_rng(seed)
    if len(scan) > n:
        idx = np.sort(rng.choice(len(scan), n, replace=False))
        return np.ascontiguousarray(scan[idx])
    extra = scan[rng.integers(0, len(scan), n - len(scan))].copy()
    extra[:, :3] += rng.normal(size=(len(extra), 3)).astype(np.float32) * 0.005
    out = np.concatenate([scan, extra])
    return np.ascontiguousarray(out[np.argsort(out[:, 3], kind="stable")])


# sensor_msgs::PointField datatype codes
PF_INT8, PF_UINT8, PF_INT16, PF_UINT16, PF_INT32, PF_UINT32, PF_FLOAT32, PF_FLOAT64 = range(1, 9)
# The message layout the reference registers for its LidarPoint (lidar/frame.hpp:12-23): x y z intensity(u8) ring(u16) timestamp(f64)
LIDAR_POINT_FIELDS = [("x", 0, PF_FLOAT32, 1), ("y", 4, PF_FLOAT32, 1), ("z", 8, PF_FLOAT32, 1), ("intensity", 12, PF_UINT8, 1),
                      ("ring", 14, PF_UINT16, 1), ("timestamp", 16, PF_FLOAT64, 1)]
LIDAR_POINT_STEP = 24


def make_pointcloud2(xyz, ring, stamp_s, intensity=None, point_step=LIDAR_POINT_STEP, fields=None):
    """Pack one scan as a sensor_msgs::PointCloud2 payload: uint8 [n, point_step] + the field list.
    stamp_s: absolute per-point firing time in seconds (float64), the `timestamp` field of the reference's LidarPoint."""
    fields = fields or LIDAR_POINT_FIELDS
    n = len(xyz)
    buf = np.zeros((n, point_step), np.uint8)
    off = {f[0]: f[1] for f in fields}
    xyz = np.ascontiguousarray(xyz, np.float32)
    for k, name in enumerate("xyz"):
        buf[:, off[name]:off[name] + 4] = xyz[:, k:k + 1].copy().view(np.uint8)
    if "intensity" in off:
        inten = np.zeros(n, np.uint8) if intensity is None else np.asarray(intensity, np.uint8)
        buf[:, off["intensity"]] = inten
    if "ring" in off:
        buf[:, off["ring"]:off["ring"] + 2] = np.asarray(ring, np.uint16).reshape(-1, 1).copy().view(np.uint8)
    if "timestamp" in off:
        buf[:, off["timestamp"]:off["timestamp"] + 8] = np.asarray(stamp_s, np.float64).reshape(-1, 1).copy().view(np.uint8)
    return buf, list(fields)
