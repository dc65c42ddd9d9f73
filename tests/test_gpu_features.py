"""GPU parity: K4 LOAM feature extraction vs the CPU oracle — index lists identical, in emission order."""
import numpy as np
import pytest

from funny_lidar_slam_b200 import synth

pytestmark = pytest.mark.gpu


def _check(proj, corner_thr=1.0, planar_thr=0.1):
    from funny_lidar_slam_b200.features import FeatureExtractor
    from oracle import pyoracle as orc
    n = len(proj["ordered"])
    fx = FeatureExtractor(corner_thr, planar_thr)
    gc, gp = fx.extract_indices(proj["depth"], proj["col"], n, proj["row_start"], proj["row_end"])
    oc, op, _ = orc.extract_features(proj["depth"], proj["col"], n, proj["row_start"], proj["row_end"], corner_thr, planar_thr)
    assert np.array_equal(gc, oc), (len(gc), len(oc))
    assert np.array_equal(gp, op), (len(gp), len(op))
    return gc, gp


def test_spinning_64_line(world, traj):
    proj = synth.make_projected_scan(world, traj[4], kind="spin", sensor="hdl64", seed=11)
    gc, gp = _check(proj)
    assert len(gc) > 100 and len(gp) > 50000


def test_spinning_16_line_thresholds(world, traj):
    proj = synth.make_projected_scan(world, traj[7], kind="spin", sensor="vlp16", seed=12)
    for ct, pt in ((1.0, 0.1), (0.2, 0.05), (5.0, 1.0)):
        _check(proj, ct, pt)


def test_livox_shaped_config3(world, traj):
    """BASELINE config 3: Livox-Avia-shaped 6 x 40000 organisation (~240k rays), thresholds 1.0 / 0.1."""
    proj = synth.make_projected_scan(world, traj[2], kind="livox", seed=13)
    gc, gp = _check(proj)
    assert len(gp) > 100000


def test_through_cluster_api(world, traj):
    from funny_lidar_slam_b200.features import FeatureExtractor
    from funny_lidar_slam_b200.registration import PointcloudCluster
    proj = synth.make_projected_scan(world, traj[1], kind="spin", sensor="vlp16", seed=14)
    cl = PointcloudCluster(ordered_cloud=proj["ordered"], point_depth_vec=proj["depth"], point_col_index_vec=proj["col"],
                           row_start_index_vec=proj["row_start"], row_end_index_vec=proj["row_end"])
    fx = FeatureExtractor(1.0, 0.1)
    fx.ExtractFeatures(cl)
    assert cl.corner_cloud.shape[1] == 4 and len(cl.planar_cloud) == len(fx.planar_idx)
    assert np.array_equal(cl.corner_cloud, proj["ordered"][fx.corner_idx])


def test_degenerate_rows():
    """rows with fewer than 12 points are skipped (block_start >= block_end), empty input returns nothing."""
    from funny_lidar_slam_b200.features import FeatureExtractor
    from oracle import pyoracle as orc
    rng = np.random.default_rng(5)
    counts = [0, 7, 11, 12, 13, 40, 3, 200]
    n = sum(counts)
    depth = rng.uniform(3, 40, n).astype(np.float32)
    col = np.concatenate([np.sort(rng.choice(1800, c, replace=False)) for c in counts]).astype(np.int32) if n else np.zeros(0, np.int32)
    ends = np.cumsum(counts)
    rs = (ends - np.array(counts) + 5).astype(np.int32)
    re = (ends - 6).astype(np.int32)
    fx = FeatureExtractor(1.0, 0.1)
    gc, gp = fx.extract_indices(depth, col, n, rs, re)
    oc, op, _ = orc.extract_features(depth, col, n, rs, re, 1.0, 0.1)
    assert np.array_equal(gc, oc) and np.array_equal(gp, op)


def _ring_case(depth_rows, col_step=1):
    counts = [len(d) for d in depth_rows]
    depth = np.concatenate(depth_rows).astype(np.float32)
    col = np.concatenate([np.arange(c) * col_step for c in counts]).astype(np.int32)
    ends = np.cumsum(counts)
    rs = (ends - np.array(counts) + 5).astype(np.int32)
    re = (ends - 6).astype(np.int32)
    return {"ordered": np.zeros((len(depth), 4), np.float32), "depth": depth, "col": col, "row_start": rs, "row_end": re}


def _noisy_rows(seed, sizes, sigma):
    rng = np.random.default_rng(seed)
    return [12.0 + 3.0 * np.sin(np.arange(c) / 90.0) + rng.normal(0, sigma, c) for c in sizes]


def test_sequential_tail_of_greedy_passes(monkeypatch):
    """The round-parallel greedy passes hand the undecided rest to one thread after `max_rounds` rounds (monotone
    roughness ramps decide one point per round); forcing that hand-over after 1 and 3 rounds must not change a bit."""
    rows = _noisy_rows(21, (300, 2500, 7000), 0.01)
    ref = _check(_ring_case(rows), 0.02, 0.004)
    assert len(ref[0]) > 40
    for rounds in ("1", "3"):
        monkeypatch.setenv("FLS_FEAT_MAX_ROUNDS", rounds)
        got = _check(_ring_case(rows), 0.02, 0.004)
        assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1])


def test_corner_limit_and_dense_corners():
    """Range noise large enough that most points qualify as corners: each block hits the 20-pick limit and the 21st
    candidate must be left untouched for the planar pass."""
    rows = _noisy_rows(8, (400, 1900, 5000), 0.03)
    gc, _ = _check(_ring_case(rows), 0.05, 0.001)
    assert len(gc) > 250  # the two long rings saturate at 6 x 20
    _check(_ring_case(rows, col_step=3), 0.05, 0.001)     # column gaps of 3: reach stays 5 (|dcol| <= 10 per step)
    _check(_ring_case(rows, col_step=11), 0.05, 0.001)    # gaps > 10: no suppression reach at all


def _raw_scan(world, pose, sensor, seed, shuffle_cols=True):
    """Raw cloud in a firing-like order (column-major over rings) with duplicate hits per cell, as a driver delivers it."""
    sc = synth.make_scan(world, pose, sensor, seed=seed)
    order = np.lexsort((sc["ring"], sc["col"]))
    pts, ring = sc["points"][order], sc["ring"][order].astype(np.int32)
    # a second, slightly different return for every 7th point: the later duplicate must lose its cell
    dup = pts[::7].copy()
    dup[:, :3] *= 1.0005
    return np.concatenate([pts, dup]), np.concatenate([ring, ring[::7]])


@pytest.mark.parametrize("sensor", ["vlp16", "hdl64"])
def test_projector_matches_oracle_bit_for_bit(world, traj, sensor):
    from funny_lidar_slam_b200.features import PointcloudProjector
    from oracle import pyoracle as orc
    sn = synth.SENSORS[sensor]
    raw, ring = _raw_scan(world, traj[3], sensor, 31)
    H, V = sn.cols, sn.lines
    h_res = float(np.float32(2 * np.pi / H))
    o = orc.project(raw, ring, V, H, h_res, 2.0, 80.0)
    g = PointcloudProjector(H, V, h_res, 2.0, 80.0).project_arrays(raw, ring)
    assert g["n"] == o["n"] and g["n"] > 1000
    n = g["n"]
    assert np.array_equal(g["ordered"], o["ordered"])
    assert np.array_equal(g["depth"][:n], o["depth"][:n]) and np.array_equal(g["col"][:n], o["col"][:n])
    assert np.array_equal(g["row_start"], o["row_start"]) and np.array_equal(g["row_end"], o["row_end"])
    # and through the pcl::PointXYZI layout, feeding the extractor (projector -> features, both on the device)
    from funny_lidar_slam_b200.features import FeatureExtractor
    from funny_lidar_slam_b200.registration import PointcloudCluster
    from tests.conftest import to_pcl
    cl = PointcloudCluster(extra=dict(raw_cloud=to_pcl(raw), ring=ring))
    PointcloudProjector(H, V, h_res, 2.0, 80.0).Project(cl)
    assert np.array_equal(cl.ordered_cloud, o["ordered"])
    fx = FeatureExtractor(1.0, 0.1)
    fx.ExtractFeatures(cl)
    oc, op, _ = orc.extract_features(o["depth"], o["col"], n, o["row_start"], o["row_end"], 1.0, 0.1)
    assert np.array_equal(fx.corner_idx, oc) and np.array_equal(fx.planar_idx, op)


def test_projector_edge_cases():
    from funny_lidar_slam_b200.features import PointcloudProjector
    from oracle import pyoracle as orc
    rng = np.random.default_rng(3)
    V, H = 4, 90
    h_res = float(np.float32(2 * np.pi / H))
    raw = np.zeros((400, 4), np.float32)
    raw[:, :3] = rng.normal(0, 12, (400, 3))
    raw[:5, :3] = 0.0            # zero range: gated out
    raw[5:10, :3] *= 100.0       # beyond max range
    ring = rng.integers(-1, V + 1, 400).astype(np.int32)  # includes invalid rings -1 and V
    o = orc.project(raw, ring, V, H, h_res, 1.0, 60.0)
    g = PointcloudProjector(H, V, h_res, 1.0, 60.0).project_arrays(raw, ring)
    assert g["n"] == o["n"]
    assert np.array_equal(g["ordered"], o["ordered"]) and np.array_equal(g["col"][:g["n"]], o["col"][:o["n"]])
    assert np.array_equal(g["row_start"], o["row_start"]) and np.array_equal(g["row_end"], o["row_end"])
    e = PointcloudProjector(H, V, h_res, 1.0, 60.0).project_arrays(np.zeros((0, 4), np.float32), np.zeros(0, np.int32))
    assert e["n"] == 0 and list(e["row_start"]) == [5] * V and list(e["row_end"]) == [-6] * V
