"""GPU parity: IcpOptimized path (K3 + K6 + K7) through the C ABI vs the CPU oracle — BASELINE config 1."""
import numpy as np
import pytest

from funny_lidar_slam_b200 import FLS_ICP_P2P, default_config, synth
from funny_lidar_slam_b200._abi import FLS_FLAG_ITER_LOG

pytestmark = pytest.mark.gpu
POS_TOL, ROT_TOL = 1e-4, 1e-4


def _pair(cfg):
    from funny_lidar_slam_b200.registration import Registration
    from oracle import pyoracle as orc
    return Registration(cfg), orc.Registration(cfg)


def _cluster(scan):
    from funny_lidar_slam_b200.registration import PointcloudCluster
    return PointcloudCluster(ordered_cloud=scan)


@pytest.mark.parametrize("guess_key", ["guess", "guess_small"])
def test_config1_localization(scene16, guess_key):
    """2 synthetic 16-line scans vs a static map, params of config/localization/config_nclt_icp.yaml:42-48."""
    cfg = default_config(FLS_ICP_P2P, flags=FLS_FLAG_ITER_LOG)
    g, o = _pair(cfg)
    g.AddCloudToLocalMap([scene16["map"]])
    o.add_cloud(scene16["map"])
    assert g.map_info().n_points == o.map_points
    Tg = scene16[guess_key].copy()
    ok_g = g.Match(_cluster(scene16["scan"]), Tg)
    ok_o, To, st_o = o.match(scene16["scan"], scene16[guess_key])
    st_g = g.last_stats
    assert ok_g == ok_o and st_g.iterations == st_o.iterations
    lg, lo = g.iter_log(), o.iter_log()
    assert lg[0]["n_valid"] == lo[0]["n_valid"]
    assert np.allclose(lg[0]["H"], lo[0]["H"], rtol=1e-9, atol=1e-6)
    assert np.allclose(lg[0]["g"], lo[0]["g"], rtol=1e-9, atol=1e-6)
    dt, dr = synth.pose_error(Tg, To)
    assert dt < POS_TOL and dr < ROT_TOL, (dt, dr)
    fo, fg = o.fitness(1.0), g.GetFitnessScore(1.0)
    assert abs(fg - fo) <= 1e-5 * max(1.0, abs(fo))
    fo2, fg2 = o.fitness(2.0), g.GetFitnessScore(2.0)  # localization.cpp:138 upstream calls it with 2.0
    assert abs(fg2 - fo2) <= 1e-5 * max(1.0, abs(fo2))


def test_too_few_points_and_no_map(scene16):
    from funny_lidar_slam_b200._lib import FlsError
    from funny_lidar_slam_b200.registration import Registration
    g = Registration(default_config(FLS_ICP_P2P))
    with pytest.raises(FlsError):
        g.Match(_cluster(scene16["scan"]), scene16["guess"].copy())  # no map yet
    g.AddCloudToLocalMap([scene16["map"]])
    with pytest.raises(FlsError):
        g.Match(_cluster(scene16["scan"][:10]), scene16["guess"].copy())  # CHECK_GT(size, 10u)


def test_mapping_mode_window(world, traj):
    cfg = default_config(FLS_ICP_P2P, localization_mode=0, local_map_size=3, dist_thre_add_cloud=0.5)
    g, o = _pair(cfg)
    first = synth.transform_points(synth.make_scan(world, traj[0], "vlp16", seed=60)["points"], traj[0])
    g.AddCloudToLocalMap([first])
    o.add_cloud(first)
    Tg_prev, To_prev = traj[0].copy(), traj[0].copy()
    for k in range(1, 5):
        scan = synth.make_scan(world, traj[k], "vlp16", seed=60 + k)["points"]
        Tg = Tg_prev.copy()
        ok_g = g.Match(_cluster(scan), Tg)
        ok_o, To, st_o = o.match(scan, To_prev)
        assert ok_g == ok_o and g.last_stats.iterations == st_o.iterations, k
        dt, dr = synth.pose_error(Tg, To)
        assert dt < POS_TOL and dr < ROT_TOL, (k, dt, dr)
        assert g.map_info().n_points == o.map_points, k
        Tg_prev, To_prev = Tg, To
