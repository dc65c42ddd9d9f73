"""Deterministic synthetic LiDAR data: analytic world + ray-cast sensors.

The reference ships no scan fixtures (datasets are external links, README.md:102-171 upstream), so
every test / bench input is generated here from fixed seeds: a square world (ground plane, 4 walls,
axis-aligned boxes, vertical cylinders), spinning 16/64/128-line sensors with the vertical geometry
of src/lidar/lidar_model.cpp:24-54 upstream, a Livox-Avia-shaped 6-line non-repetitive pattern, a
figure-8 trajectory with ground-truth poses, and map builders.  Pure numpy; shared by tests, bench.py
and __graft_entry__.smoke().  This is input synthesis, not part of the registration path.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

__all__ = [
    "World", "make_world", "raycast", "Sensor", "SENSORS", "sensor_dirs", "make_scan", "trajectory", "perturb_pose",
    "voxel_downsample_np", "make_map_from_scans", "make_surface_map", "livox_dirs", "make_projected_scan", "se3", "pose_error",
]


@dataclass
class World:
    half: float          # ground / walls span [-half, half]^2
    wall_h: float
    boxes: np.ndarray    # (nb, 6) cx, cy, cz, hx, hy, hz
    cyls: np.ndarray     # (nc, 4) cx, cy, r, h


def make_world(seed: int = 1234, half: float = 100.0, n_boxes: int = 40, n_cyls: int = 30, keepout: float = 6.0) -> World:
    rng = np.random.default_rng(seed)
    sz = rng.uniform(1.0, 5.0, size=(n_boxes, 3))            # half extents: 2..10 m boxes
    sz[:, 2] = rng.uniform(0.75, 4.0, size=n_boxes)
    cxy = rng.uniform(-half + 8, half - 8, size=(n_boxes, 2))
    # keep the trajectory corridor (figure-8 around the origin) free of box interiors
    near = np.linalg.norm(cxy, axis=1) < keepout + np.linalg.norm(sz[:, :2], axis=1)
    cxy[near] += np.sign(cxy[near] + 1e-9) * (keepout + 8.0)
    boxes = np.concatenate([cxy, sz[:, 2:3], sz], axis=1)
    cyl = np.stack([rng.uniform(-half + 5, half - 5, n_cyls), rng.uniform(-half + 5, half - 5, n_cyls),
                    rng.uniform(0.15, 0.5, n_cyls), rng.uniform(3.0, 9.0, n_cyls)], axis=1)
    return World(half=float(half), wall_h=12.0, boxes=boxes.astype(np.float64), cyls=cyl.astype(np.float64))


def raycast(world: World, origin: np.ndarray, dirs: np.ndarray, max_range: float = 100.0, chunk: int = 32768):
    """Nearest positive hit distance along unit rays from `origin`; returns (t, surface_id)."""
    o = np.asarray(origin, np.float64)
    n = dirs.shape[0]
    t_out = np.full(n, np.inf)
    sid = np.zeros(n, np.int32)
    # cull primitives beyond reach
    b = world.boxes
    if len(b):
        reach = max_range + np.linalg.norm(b[:, 3:6], axis=1)
        b = b[np.linalg.norm(b[:, :3] - o, axis=1) < reach]
    c = world.cyls
    if len(c):
        c = c[np.hypot(c[:, 0] - o[0], c[:, 1] - o[1]) < max_range + c[:, 2]]
    H = world.half
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        for s in range(0, n, chunk):
            d = dirs[s:s + chunk].astype(np.float64)
            m = d.shape[0]
            best = np.full(m, np.inf)
            bid = np.zeros(m, np.int32)
            # ground z = 0
            tg = -o[2] / d[:, 2]
            ok = (tg > 1e-6)
            xg = o[0] + tg * d[:, 0]
            yg = o[1] + tg * d[:, 1]
            ok &= (np.abs(xg) <= H) & (np.abs(yg) <= H)
            tg = np.where(ok, tg, np.inf)
            bid = np.where(tg < best, 1, bid)
            best = np.minimum(best, tg)
            # walls
            for axis, sign, wid in ((0, 1, 2), (0, -1, 3), (1, 1, 4), (1, -1, 5)):
                tw = (sign * H - o[axis]) / d[:, axis]
                other = 1 - axis
                po = o[other] + tw * d[:, other]
                pz = o[2] + tw * d[:, 2]
                ok = (tw > 1e-6) & (np.abs(po) <= H) & (pz >= 0) & (pz <= world.wall_h)
                tw = np.where(ok, tw, np.inf)
                bid = np.where(tw < best, wid, bid)
                best = np.minimum(best, tw)
            # boxes (slab test), rays x boxes
            if len(b):
                inv = 1.0 / d  # (m,3)
                lo = (b[None, :, :3] - b[None, :, 3:6] - o[None, None, :]) * inv[:, None, :]
                hi = (b[None, :, :3] + b[None, :, 3:6] - o[None, None, :]) * inv[:, None, :]
                tn = np.minimum(lo, hi).max(axis=2)
                tf = np.maximum(lo, hi).min(axis=2)
                tb = np.where((tf >= tn) & (tn > 1e-6), tn, np.inf)
                k = tb.argmin(axis=1)
                tbm = tb[np.arange(m), k]
                bid = np.where(tbm < best, 10 + k.astype(np.int32), bid)
                best = np.minimum(best, tbm)
            # vertical cylinders: |(o+td)_xy - c| = r, 0 <= z <= h
            if len(c):
                ox = o[0] - c[None, :, 0]
                oy = o[1] - c[None, :, 1]
                A = (d[:, 0] ** 2 + d[:, 1] ** 2)[:, None]
                B = 2 * (ox * d[:, 0:1] + oy * d[:, 1:2])
                C = ox ** 2 + oy ** 2 - c[None, :, 2] ** 2
                disc = B * B - 4 * A * C
                sq = np.sqrt(np.where(disc > 0, disc, np.nan))
                tc = (-B - sq) / (2 * A)
                zc = o[2] + tc * d[:, 2:3]
                tc = np.where((tc > 1e-6) & (zc >= 0) & (zc <= c[None, :, 3]), tc, np.inf)
                k = tc.argmin(axis=1)
                tcm = tc[np.arange(m), k]
                bid = np.where(tcm < best, 5000 + k.astype(np.int32), bid)
                best = np.minimum(best, tcm)
            t_out[s:s + chunk] = best
            sid[s:s + chunk] = bid
    return t_out, sid


@dataclass(frozen=True)
class Sensor:
    name: str
    lines: int
    cols: int
    elev_lo_deg: float
    elev_hi_deg: float


# vertical geometry follows src/lidar/lidar_model.cpp:24-54 upstream (lower angle / vertical resolution)
SENSORS = {
    "vlp16": Sensor("Velodyne_16", 16, 1800, -15.0, 15.0),
    "hdl64": Sensor("Velodyne_64", 64, 1800, -24.9, 2.0),
    "os128": Sensor("Ouster_128", 128, 2048, -22.5, 22.5),
}


def sensor_dirs(sensor: Sensor):
    """Unit directions in the body frame, firing order = column-major (all rings of a column, then next column)."""
    elev = np.deg2rad(np.linspace(sensor.elev_lo_deg, sensor.elev_hi_deg, sensor.lines))
    az = -np.pi + 2 * np.pi * (np.arange(sensor.cols) + 0.5) / sensor.cols
    ring = np.tile(np.arange(sensor.lines, dtype=np.int32), sensor.cols)
    col = np.repeat(np.arange(sensor.cols, dtype=np.int32), sensor.lines)
    ce = np.cos(elev)[ring]
    d = np.stack([ce * np.cos(az[col]), ce * np.sin(az[col]), np.sin(elev)[ring]], axis=1)
    return d, ring, col


def livox_dirs(n_lines: int = 6, n_samples: int = 40000, fov_h_deg: float = 70.4, fov_v_deg: float = 77.2):
    """Livox-Avia-shaped non-repetitive rosette: 6 lines x n_samples, line-major (row = line)."""
    k = np.arange(n_samples, dtype=np.float64)
    ring = np.repeat(np.arange(n_lines, dtype=np.int32), n_samples)
    col = np.tile(np.arange(n_samples, dtype=np.int32), n_lines)
    ph = (2 * np.pi / n_lines) * ring
    tt = k[col] / n_samples
    # two incommensurate rotations (Risley prisms) -> rosette
    a = 2 * np.pi * 7.0 * tt + ph
    bq = 2 * np.pi * (-4.6180339887) * tt + 0.37 * ph
    u = 0.5 * (np.cos(a) + np.cos(bq))
    v = 0.5 * (np.sin(a) + np.sin(bq))
    az = np.deg2rad(fov_h_deg / 2) * u
    el = np.deg2rad(fov_v_deg / 2) * v
    d = np.stack([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el)], axis=1)
    return d, ring, col


def se3(xyz, rpy) -> np.ndarray:
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    T = np.eye(4)
    T[:3, :3] = Rz @ Ry @ Rx
    T[:3, 3] = xyz
    return T


def trajectory(n: int, step: float = 1.0, scale: float = 30.0, height: float = 1.8) -> np.ndarray:
    """Figure-8 (lemniscate-like) ground-truth body poses, ~`step` metres apart; (n, 4, 4)."""
    s = np.linspace(0, 2 * np.pi, 20000)
    x = scale * np.sin(s)
    y = scale * 0.5 * np.sin(2 * s)
    arc = np.concatenate([[0], np.cumsum(np.hypot(np.diff(x), np.diff(y)))])
    want = np.arange(n) * step
    want = np.mod(want, arc[-1])
    si = np.interp(want, arc, s)
    xs, ys = scale * np.sin(si), scale * 0.5 * np.sin(2 * si)
    yaw = np.arctan2(scale * np.cos(2 * si), scale * np.cos(si))
    out = np.empty((n, 4, 4))
    for i in range(n):
        out[i] = se3([xs[i], ys[i], height], [0.01 * np.sin(3 * si[i]), 0.015 * np.cos(2 * si[i]), yaw[i]])
    return out


def perturb_pose(T: np.ndarray, seed: int = 77, dpos: float = 0.3, drot_deg: float = 3.0) -> np.ndarray:
    rng = np.random.default_rng(seed)
    v = rng.normal(size=3)
    v *= dpos / np.linalg.norm(v)
    w = rng.normal(size=3)
    w *= np.deg2rad(drot_deg) / np.linalg.norm(w)
    th = np.linalg.norm(w)
    k = w / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    dR = np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K
    out = T.copy()
    out[:3, :3] = dR @ T[:3, :3]
    out[:3, 3] = T[:3, 3] + v
    return out


def pose_error(Ta: np.ndarray, Tb: np.ndarray):
    """(translation error [m], rotation angle [rad]) between two 4x4 poses."""
    dt = float(np.linalg.norm(Ta[:3, 3] - Tb[:3, 3]))
    dR = Ta[:3, :3].T @ Tb[:3, :3]
    c = np.clip((np.trace(dR) - 1) / 2, -1, 1)
    s = np.linalg.norm([dR[2, 1] - dR[1, 2], dR[0, 2] - dR[2, 0], dR[1, 0] - dR[0, 1]]) / 2
    return dt, float(np.arctan2(s, c))


def make_scan(world: World, pose: np.ndarray, sensor="hdl64", seed: int = 0, noise: float = 0.02, min_range: float = 2.0,
              max_range: float = 100.0, dirs=None):
    """Ray-cast one scan.  Returns dict(points (n,4) float32 body-frame xyz+intensity, ring, col, depth)."""
    if dirs is None:
        sensor = SENSORS[sensor] if isinstance(sensor, str) else sensor
        d, ring, col = sensor_dirs(sensor)
    else:
        d, ring, col = dirs
    R, t = pose[:3, :3], pose[:3, 3]
    dw = d @ R.T
    tt, sid = raycast(world, t, dw, max_range)
    rng = np.random.default_rng(1000003 * (seed + 1))
    tt = tt + rng.normal(0.0, noise, size=tt.shape)
    keep = np.isfinite(tt) & (tt >= min_range) & (tt <= max_range)
    pts = (d[keep] * tt[keep, None]).astype(np.float32)
    inten = ((sid[keep] % 7) * 10.0 + 5.0).astype(np.float32)
    out = np.concatenate([pts, inten[:, None]], axis=1).astype(np.float32)
    return {"points": np.ascontiguousarray(out), "ring": ring[keep].copy(), "col": col[keep].copy(),
            "depth": np.linalg.norm(pts, axis=1).astype(np.float32)}


def transform_points(points: np.ndarray, T: np.ndarray) -> np.ndarray:
    out = points.copy()
    out[:, :3] = (points[:, :3].astype(np.float64) @ T[:3, :3].T + T[:3, 3]).astype(np.float32)
    return out


def voxel_downsample_np(points: np.ndarray, leaf: float) -> np.ndarray:
    """Centroid voxel filter for DATA SYNTHESIS only (fp64 accumulation; not the PCL restatement)."""
    if len(points) == 0:
        return points
    key = np.floor(points[:, :3].astype(np.float64) / leaf).astype(np.int64)
    key -= key.min(axis=0)
    dims = key.max(axis=0) + 1
    lin = key[:, 0] + dims[0] * (key[:, 1] + dims[1] * key[:, 2])
    uniq, inv, cnt = np.unique(lin, return_inverse=True, return_counts=True)
    out = np.zeros((len(uniq), 4))
    for a in range(4):
        out[:, a] = np.bincount(inv, weights=points[:, a].astype(np.float64), minlength=len(uniq)) / cnt
    return np.ascontiguousarray(out.astype(np.float32))


def make_map_from_scans(world: World, poses, sensor="hdl64", leaf: float = 0.3, seed0: int = 5000, **kw) -> np.ndarray:
    """Union of scans taken at ground-truth poses, expressed in the map frame and voxel-filtered
    (what Localization::LoadLocalMap feeds AddCloudToLocalMap upstream, src/slam/localization.cpp:180-181)."""
    clouds = []
    for i, T in enumerate(poses):
        sc = make_scan(world, T, sensor, seed=seed0 + i, **kw)
        clouds.append(transform_points(sc["points"], T))
    return voxel_downsample_np(np.concatenate(clouds, axis=0), leaf)


def make_surface_map(world: World, spacing: float = 0.3, seed: int = 4321, noise: float = 0.02, max_points: int | None = None) -> np.ndarray:
    """Directly sample every surface of the world on a jittered grid (a voxel-filtered prior map without
    ray casting).  Used for the multi-million-point iVox map of BASELINE config 4.  Written to touch little
    fresh memory: one float32 output array filled strip by strip."""
    rng = np.random.default_rng(seed)
    H = world.half
    # surfaces as (u0, u1, v0, v1, kind, params)
    surf = [(-H, H, -H, H, "z", (0.0,))]
    for sgn in (-1.0, 1.0):
        surf.append((-H, H, 0, world.wall_h, "x", (sgn * H,)))
        surf.append((-H, H, 0, world.wall_h, "y", (sgn * H,)))
    for cx, cy, cz, hx, hy, hz in world.boxes:
        surf.append((cx - hx, cx + hx, cy - hy, cy + hy, "z", (cz + hz,)))
        for s in (-1.0, 1.0):
            surf.append((cy - hy, cy + hy, cz - hz, cz + hz, "x", (cx + s * hx,)))
            surf.append((cx - hx, cx + hx, cz - hz, cz + hz, "y", (cy + s * hy,)))
    for cx, cy, r, h in world.cyls:
        surf.append((0.0, 2 * np.pi * r, 0.0, h, "c", (cx, cy, r)))
    dims = [(max(1, int(round((u1 - u0) / spacing))), max(1, int(round((v1 - v0) / spacing)))) for u0, u1, v0, v1, _, _ in surf]
    total = sum(a * b for a, b in dims)
    out = np.empty((total, 4), np.float32)
    out[:, 3] = 20.0
    pos = 0
    strip = 1 << 20
    for (u0, u1, v0, v1, kind, prm), (nu, nv) in zip(surf, dims):
        m = nu * nv
        du, dv = (u1 - u0) / nu, (v1 - v0) / nv
        for s0 in range(0, m, strip):
            k = np.arange(s0, min(m, s0 + strip))
            uu = (u0 + ((k // nv) + 0.5) * du + rng.uniform(-0.45, 0.45, k.size) * spacing).astype(np.float32)
            vv = (v0 + ((k % nv) + 0.5) * dv + rng.uniform(-0.45, 0.45, k.size) * spacing).astype(np.float32)
            ee = rng.normal(0, noise, k.size).astype(np.float32)
            o = out[pos + s0: pos + s0 + k.size]
            if kind == "z":
                o[:, 0], o[:, 1], o[:, 2] = uu, vv, prm[0] + ee
            elif kind == "x":
                o[:, 0], o[:, 1], o[:, 2] = prm[0] + ee, uu, vv
            elif kind == "y":
                o[:, 0], o[:, 1], o[:, 2] = uu, prm[0] + ee, vv
            else:
                cx, cy, r = prm
                ang = uu / np.float32(r)
                o[:, 0], o[:, 1], o[:, 2] = cx + (r + ee) * np.cos(ang), cy + (r + ee) * np.sin(ang), vv
        pos += m
    if max_points is not None and len(out) > max_points:
        out = out[np.sort(rng.permutation(len(out))[:max_points])]
    return out


def make_projected_scan(world: World, pose: np.ndarray, kind: str = "livox", seed: int = 3, lines: int = 6, samples: int = 40000,
                        sensor: str = "hdl64", noise: float = 0.02, min_range: float = 2.0, max_range: float = 100.0):
    """Projector-format arrays (PointcloudCluster::ordered_cloud_ / point_depth_vec_ / point_col_index_vec_ /
    row_start_index_vec_ / row_end_index_vec_, src/loam/pointcloud_projector.cpp:113-132 upstream) for a
    rows x cols organised scan: row-major compaction of valid returns, row_start = first+5, row_end = last-5."""
    if kind == "livox":
        d, ring, col = livox_dirs(lines, samples)
        V = lines
    else:
        sn = SENSORS[sensor]
        d, ring, col = sensor_dirs(sn)
        V = sn.lines
    sc = make_scan(world, pose, seed=seed, noise=noise, min_range=min_range, max_range=max_range, dirs=(d, ring, col))
    order = np.lexsort((sc["col"], sc["ring"]))
    pts = sc["points"][order]
    rg = sc["ring"][order]
    cl = sc["col"][order]
    depth = sc["depth"][order]
    counts = np.bincount(rg, minlength=V)
    ends = np.cumsum(counts)
    starts = ends - counts
    return {"ordered": np.ascontiguousarray(pts), "depth": np.ascontiguousarray(depth), "col": np.ascontiguousarray(cl.astype(np.int32)),
            "row_start": (starts + 5).astype(np.int32), "row_end": (ends - 6).astype(np.int32), "rows": V}
