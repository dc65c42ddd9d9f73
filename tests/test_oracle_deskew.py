"""The oracle's pre-hot-path pipeline (oracle/orc_deskew.h: LidarDistortionCorrector + PreProcessing::Run's non-feature branch +
the projector's de-skew hook) against an independent numpy restatement of the cited reference lines."""
import numpy as np

from oracle import pyoracle as orc


def make_case(n=6000, seed=0, t_end_imu=0.085):
    rng = np.random.default_rng(seed)
    raw = np.zeros((n, 5), np.float32)
    raw[:, :3] = rng.normal(0, 12, (n, 3))
    raw[:, 3] = rng.random(n)
    raw[:, 4] = np.linspace(0.0, 0.1, n)  # sweep time relative to the reference
    ref = 1_700_000_000_000_000
    t = (ref - 20_000 + np.arange(0, int((0.02 + t_end_imu) * 1e6), 5000)).astype(np.uint64)  # 200 Hz, ends before the sweep does
    ang = 0.3 * np.sin(np.linspace(0, 2.0, len(t)))
    axis = np.array([0.2, -0.1, 0.97]) / np.linalg.norm([0.2, -0.1, 0.97])
    q = np.concatenate([axis[None, :] * np.sin(ang / 2)[:, None], np.cos(ang / 2)[:, None]], 1)
    T = np.eye(4)
    T[:3, :3] = np.array([[0.0, -1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]])
    T[:3, 3] = [0.05, -0.02, 0.1]
    return raw, dict(t_us=t, q_xyzw=q, ref_time_us=ref, T_lidar_to_imu=T)


def np_quat_mul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz, aw * bz + az * bw + ax * by - ay * bx,
                     aw * bw - ax * bx - ay * by - az * bz])


def np_rot(q, v):
    R = np.array([[1 - 2 * (q[1] ** 2 + q[2] ** 2), 2 * (q[0] * q[1] - q[2] * q[3]), 2 * (q[0] * q[2] + q[1] * q[3])],
                  [2 * (q[0] * q[1] + q[2] * q[3]), 1 - 2 * (q[0] ** 2 + q[2] ** 2), 2 * (q[1] * q[2] - q[0] * q[3])],
                  [2 * (q[0] * q[2] - q[1] * q[3]), 2 * (q[1] * q[2] + q[0] * q[3]), 1 - 2 * (q[0] ** 2 + q[1] ** 2)]])
    return R @ v


def np_interp(imu, t):
    ts = imu["t_us"].astype(np.int64)
    if t < ts[0] or t > ts[-1]:
        return None
    r = int(np.searchsorted(ts, t, side="right"))
    r = min(max(r, 1), len(ts) - 1)
    l = r - 1
    s = (t - ts[l]) / float(ts[r] - ts[l])
    q = imu["q_xyzw"][l] * (1 - s) + imu["q_xyzw"][r] * s
    return q / np.linalg.norm(q)


def np_preprocess(raw, imu, min_d, max_d, jump):
    q_ref = np_interp(imu, int(imu["ref_time_us"]))
    q_ref_inv = q_ref * np.array([-1, -1, -1, 1.0])
    T = imu["T_lidar_to_imu"]
    ordered, planar_idx = [], []
    for i, p in enumerate(raw):
        d = np.float32(np.sqrt(np.float32(p[0] * p[0] + p[1] * p[1] + p[2] * p[2])))
        if d < min_d or d > max_d:
            continue
        t = int(imu["ref_time_us"]) + int(np.float64(p[4]) * 1.0e6)
        qc = np_interp(imu, t)
        if qc is None:
            continue
        v = T[:3, :3] @ p[:3].astype(np.float64) + T[:3, 3]
        c = np_rot(np_quat_mul(q_ref_inv, qc), v)
        if i % jump == 0:
            planar_idx.append(len(ordered))
        ordered.append([c[0], c[1], c[2], p[3]])
    return np.array(ordered, np.float32), planar_idx


def test_preprocess_matches_numpy_restatement():
    raw, imu = make_case()
    o_ord, o_pl = orc.preprocess(raw, imu, 2.0, 60.0, 4, 0.5)
    n_ord, pl_idx = np_preprocess(raw, imu, np.float32(2.0), np.float32(60.0), 4)
    assert len(o_ord) == len(n_ord) and 0 < len(o_ord) < len(raw)  # the range gate and the end of the IMU buffer both drop points
    assert np.abs(o_ord - n_ord).max() < 2e-5  # fp64 algebra rounded once to float vs a rotation-matrix formulation
    ref_pl = orc.voxel_grid(o_ord[pl_idx], 0.5)
    assert len(o_pl) == len(ref_pl) and np.array_equal(o_pl, ref_pl)
    # without an IMU buffer: pure gate + jump span + voxel filter
    o2, p2 = orc.preprocess(raw, None, 2.0, 60.0, 4, 0.5)
    d = np.sqrt((raw[:, :3].astype(np.float32) ** 2).sum(1, dtype=np.float32))
    keep = (d >= 2.0) & (d <= 60.0)
    assert np.array_equal(o2, raw[keep][:, :4])


def test_ref_time_outside_buffer_drops_the_scan():
    raw, imu = make_case(n=500)
    imu = dict(imu, ref_time_us=int(imu["t_us"][-1]) + 10)
    o_ord, o_pl = orc.preprocess(raw, imu, 2.0, 60.0, 4, 0.5)
    assert len(o_ord) == 0 and len(o_pl) == 0


def test_project_imu_matches_plain_projection_plus_deskew():
    """The projector keeps the first in-range point of every cell WHOSE TIME IS INSIDE THE IMU BUFFER and de-skews it; depth stays
    the raw range (pointcloud_projector.cpp:57,100-105)."""
    raw5, imu = make_case(n=20000, seed=3)
    rng = np.random.default_rng(1)
    ring = rng.integers(0, 16, len(raw5)).astype(np.int32)
    V, H, h_res = 16, 900, np.float32(2 * np.pi / 900)
    out = orc.project_imu(raw5[:, :4], ring, raw5[:, 4], imu, V, H, h_res, 2.0, 60.0)
    # restatement: drop the points ProcessPoint rejects, project the rest, then de-skew the winners
    ts = imu["t_us"].astype(np.int64)
    t = int(imu["ref_time_us"]) + (raw5[:, 4].astype(np.float64) * 1.0e6).astype(np.int64)
    ok = (t >= ts[0]) & (t <= ts[-1])
    plain = orc.project(raw5[ok][:, :4], ring[ok], V, H, h_res, 2.0, 60.0)
    assert out["n"] == plain["n"] and out["n"] > 1000
    assert np.array_equal(out["depth"], plain["depth"]) and np.array_equal(out["col"], plain["col"])
    assert np.array_equal(out["row_start"], plain["row_start"]) and np.array_equal(out["row_end"], plain["row_end"])
    # every emitted point is the de-skewed version of the plain winner
    raw_of = {tuple(p[:3]): p for p in raw5}
    q_ref = np_interp(imu, int(imu["ref_time_us"]))
    q_ref_inv = q_ref * np.array([-1, -1, -1, 1.0])
    T = imu["T_lidar_to_imu"]
    for k in range(0, out["n"], 97):
        p = raw_of[tuple(plain["ordered"][k][:3])]
        qc = np_interp(imu, int(imu["ref_time_us"]) + int(np.float64(p[4]) * 1.0e6))
        c = np_rot(np_quat_mul(q_ref_inv, qc), T[:3, :3] @ p[:3].astype(np.float64) + T[:3, 3])
        assert np.abs(out["ordered"][k][:3] - c).max() < 2e-5


def test_against_round2_goldens(world, traj):
    """Regression anchors of the round-2 oracle parts (tests/golden/make_golden.py round2()): they pin the ORACLE, not the reference."""
    import os

    from funny_lidar_slam_b200 import synth
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    g = np.load(os.path.join(gold, "deskew_case0.npz"))
    raw, imu = make_case(n=6000, seed=0)
    o_ord, o_pl = orc.preprocess(raw, imu, 2.0, 60.0, 4, 0.5)
    assert len(o_ord) == int(g["n_ordered"]) and len(o_pl) == int(g["n_planar"])
    assert np.array_equal(o_ord[:16], g["ordered_first"]) and np.array_equal(o_ord[-16:], g["ordered_last"])
    assert float(np.sum(o_ord.astype(np.float64))) == float(g["ordered_checksum"])
    assert float(np.sum(o_pl.astype(np.float64))) == float(g["planar_checksum"])
    ring = np.random.default_rng(1).integers(0, 16, len(raw)).astype(np.int32)
    pr = orc.project_imu(raw[:, :4], ring, raw[:, 4], imu, 16, 900, np.float32(2 * np.pi / 900), 2.0, 60.0)
    assert pr["n"] == int(g["proj_n"]) and float(np.sum(pr["ordered"].astype(np.float64))) == float(g["proj_checksum"])
    assert float(np.sum(pr["depth"].astype(np.float64))) == float(g["proj_depth_checksum"]) and int(np.sum(pr["col"].astype(np.int64))) == int(g["proj_col_checksum"])
    assert np.array_equal(pr["row_start"], g["proj_row_start"]) and np.array_equal(pr["row_end"], g["proj_row_end"])
    lru = np.load(os.path.join(gold, "ivox_lru_cap5000.npz"))["counts"]
    iv = orc.IVox(0.5, 2, 5000)
    for k in range(3):
        iv.add(synth.transform_points(synth.make_scan(world, traj[k], "vlp16", seed=60 + k)["points"], traj[k]))
        assert (iv.num_voxels, iv.num_points) == tuple(lru[k]), k
    assert lru[1][0] == 4999 and lru[0][0] < 4999  # the capacity is crossed by the second cloud
