"""GPU parity: the kd-tree LOAM plug-ins (K5) through the C ABI vs the CPU oracle.

LoamPointToPlaneKdtree (loam_point_to_plane_kdtree.h upstream) and LoamFull (loam_full_kdtree.h): exact unbounded
5-NN, plane / line residuals, stale-flag summation, key-frame gated sliding-window maps."""
import numpy as np
import pytest

from funny_lidar_slam_b200 import default_config, synth
from funny_lidar_slam_b200._abi import FLS_FLAG_ITER_LOG, FLS_LOAM_FULL, FLS_P2PLANE_KNN

pytestmark = pytest.mark.gpu
POS_TOL, ROT_TOL = 1e-4, 1e-4


def _pair(cfg):
    from funny_lidar_slam_b200.registration import Registration
    from oracle import pyoracle as orc
    return Registration(cfg), orc.Registration(cfg)


def _features(world, pose, seed, sensor="vlp16"):
    """(planar, corner) body-frame feature clouds through the oracle's extractor (the GPU extractor is tested elsewhere)."""
    from oracle import pyoracle as orc
    proj = synth.make_projected_scan(world, pose, kind="spin", sensor=sensor, seed=seed)
    n = len(proj["ordered"])
    ci, pi, _ = orc.extract_features(proj["depth"], proj["col"], n, proj["row_start"], proj["row_end"], 1.0, 0.1)
    return proj["ordered"][pi].copy(), proj["ordered"][ci].copy()


def _to_world(pts, T):
    out = pts.copy()
    out[:, :3] = (pts[:, :3].astype(np.float64) @ T[:3, :3].T + T[:3, 3]).astype(np.float32)
    return out


def _cluster(planar, corner=None):
    from funny_lidar_slam_b200.registration import PointcloudCluster
    return PointcloudCluster(planar_cloud=planar, corner_cloud=corner)


def _same_map_size(g, o, k):
    """Key-frame clouds enter the map transformed by the estimated pose; GPU and oracle poses agree to ~1e-6 m, so a
    point sitting on a voxel face can land on either side: sizes agree to a few points, not necessarily exactly."""
    a, b = g.map_info().n_points, o.map_points
    assert abs(a - b) <= max(2, b // 2000), (k, a, b)


def _compare_logs(lg, lo, h_rtol=1e-7):
    assert len(lg) == len(lo)
    for a, b in zip(lg, lo):
        assert a["n_valid"] == b["n_valid"]
        scale = np.abs(b["H"]).max()
        assert np.allclose(a["H"], b["H"], rtol=h_rtol, atol=1e-9 * scale)
        assert np.allclose(a["g"], b["g"], rtol=h_rtol, atol=1e-9 * max(1.0, np.abs(b["g"]).max()))


@pytest.fixture(scope="module")
def feature_scene(world, traj):
    maps_p, maps_c = [], []
    for k in (3, 4, 6, 7):
        p, c = _features(world, traj[k], k)
        maps_p.append(_to_world(p, traj[k]))
        maps_c.append(_to_world(c, traj[k]))
    p5, c5 = _features(world, traj[5], 55)
    return dict(maps_p=maps_p, maps_c=maps_c, planar=p5, corner=c5, truth=traj[5],
                guess=synth.perturb_pose(traj[5], dpos=0.1, drot_deg=1.0), guess_big=synth.perturb_pose(traj[5]))


@pytest.mark.parametrize("guess_key", ["guess", "guess_big"])
def test_point_to_plane_kdtree_localization(feature_scene, guess_key):
    cfg = default_config(FLS_P2PLANE_KNN, flags=FLS_FLAG_ITER_LOG)
    g, o = _pair(cfg)
    mp = np.concatenate(feature_scene["maps_p"])
    g.AddCloudToLocalMap([mp])
    o.add_cloud(mp)
    assert g.map_info().n_points == o.map_points
    Tg = feature_scene[guess_key].copy()
    ok_g = g.Match(_cluster(feature_scene["planar"]), Tg)
    ok_o, To, st_o = o.match(feature_scene["planar"], feature_scene[guess_key])
    st_g = g.last_stats
    assert ok_g == ok_o and st_g.iterations == st_o.iterations and st_g.n_valid == st_o.n_valid
    _compare_logs(g.iter_log(), o.iter_log())
    dt, dr = synth.pose_error(Tg, To)
    assert dt < POS_TOL and dr < ROT_TOL, (dt, dr)
    if guess_key == "guess":
        et, er = synth.pose_error(Tg, feature_scene["truth"])
        assert et < 0.05 and er < 0.01
    for rng in (1.0, 2.0):
        fo, fg = o.fitness(rng), g.GetFitnessScore(rng)
        assert abs(fg - fo) <= 1e-5 * max(1.0, abs(fo))


def test_point_to_plane_kdtree_far_guess_is_exact_and_fails_alike(feature_scene):
    """A guess hundreds of metres away: every query leaves the grid's ring search and takes the exhaustive scan; the
    5-NN (hence H, g, n_valid of iteration 0) must still be the kd-tree's."""
    cfg = default_config(FLS_P2PLANE_KNN, flags=FLS_FLAG_ITER_LOG, max_iterations=1)
    g, o = _pair(cfg)
    mp = np.concatenate(feature_scene["maps_p"][:2])[::7]
    g.AddCloudToLocalMap([mp])
    o.add_cloud(mp)
    far = feature_scene["guess"].copy()
    far[:3, 3] += (300.0, -250.0, 40.0)
    sub = feature_scene["planar"][::40]
    Tg = far.copy()
    ok_g = g.Match(_cluster(sub), Tg)
    ok_o, To, _ = o.match(sub, far)
    assert ok_g == ok_o
    lg, lo = g.iter_log(), o.iter_log()
    assert lg[0]["n_valid"] == lo[0]["n_valid"]
    scale = max(1.0, np.abs(lo[0]["H"]).max())
    assert np.allclose(lg[0]["H"], lo[0]["H"], rtol=1e-6, atol=1e-9 * scale)


def test_point_to_plane_kdtree_mapping_stream(world, traj):
    """Mapping mode: key-frame gated insertion (static last_T [quirk 7]), sliding window of 3 clouds, VoxelGrid + rebuild."""
    cfg = default_config(FLS_P2PLANE_KNN, localization_mode=0, local_map_size=3, dist_thre_add_cloud=0.5)
    g, o = _pair(cfg)
    first = None
    for k0 in (0, 2, 4):  # a well-conditioned start: three surrounding key-frames (window = 3)
        p0, _ = _features(world, traj[k0], 100 + k0)
        first = _to_world(p0, traj[k0])
        g.AddCloudToLocalMap([first])
        o.add_cloud(first)
    for k in range(1, 7):
        pk, _ = _features(world, traj[k], 100 + k)
        guess_g = synth.perturb_pose(traj[k], dpos=0.05, drot_deg=0.5, seed=k)
        Tg = guess_g.copy()
        ok_g = g.Match(_cluster(pk), Tg)
        ok_o, To, st_o = o.match(pk, guess_g)
        assert ok_g == ok_o and g.last_stats.iterations == st_o.iterations
        dt, dr = synth.pose_error(Tg, To)
        assert dt < POS_TOL and dr < ROT_TOL, (k, dt, dr)
        _same_map_size(g, o, k)
        fo, fg = o.fitness(1.0), g.GetFitnessScore(1.0)
        assert abs(fg - fo) <= 1e-4 * max(1.0, abs(fo)), k


def test_loam_full_match(feature_scene):
    cfg = default_config(FLS_LOAM_FULL, localization_mode=0, flags=FLS_FLAG_ITER_LOG)
    g, o = _pair(cfg)
    for mp, mc in zip(feature_scene["maps_p"], feature_scene["maps_c"]):
        g.AddCloudToLocalMap([mp, mc])
        o.add_cloud(mp, mc)
    assert g.map_info().n_points == o.map_points
    Tg = feature_scene["guess"].copy()
    ok_g = g.Match(_cluster(feature_scene["planar"], feature_scene["corner"]), Tg)
    ok_o, To, st_o = o.match(feature_scene["planar"], feature_scene["guess"], corner=feature_scene["corner"])
    st_g = g.last_stats
    assert ok_g and ok_o and st_g.iterations == st_o.iterations and st_g.n_valid == st_o.n_valid
    _compare_logs(g.iter_log(), o.iter_log())
    dt, dr = synth.pose_error(Tg, To)
    assert dt < POS_TOL and dr < ROT_TOL, (dt, dr)
    et, er = synth.pose_error(Tg, feature_scene["truth"])
    assert et < 0.05 and er < 0.01
    assert g.GetFitnessScore(1.0) == pytest.approx(float(np.finfo(np.float32).max))  # FloatNaN (loam_full_kdtree.h:206-208)


def test_loam_full_needs_two_clouds(feature_scene):
    from funny_lidar_slam_b200._lib import FlsError
    from funny_lidar_slam_b200.registration import Registration
    g = Registration(default_config(FLS_LOAM_FULL))
    with pytest.raises(FlsError):
        g.AddCloudToLocalMap([feature_scene["maps_p"][0]])  # CHECK_EQ(cloud_list.size(), 2)
    with pytest.raises(FlsError):
        g.Match(_cluster(feature_scene["planar"], feature_scene["corner"]), feature_scene["guess"].copy())  # no map


def test_loam_full_stream_with_filter_threshold(world, traj):
    """Eight key-frames: the maps stay unfiltered up to 5 clouds and are voxel-filtered from the 6th on
    (loam_full_kdtree.h:91-99); windows of 6 planar / 4 corner clouds slide."""
    cfg = default_config(FLS_LOAM_FULL, localization_mode=0, local_map_size=6, corner_local_map_size=4, dist_thre_add_cloud=0.5)
    g, o = _pair(cfg)
    for k0 in (0, 2, 4, 6):  # a well-conditioned start: four surrounding key-frames
        p0, c0 = _features(world, traj[k0], 200 + k0)
        g.AddCloudToLocalMap([_to_world(p0, traj[k0]), _to_world(c0, traj[k0])])
        o.add_cloud(_to_world(p0, traj[k0]), _to_world(c0, traj[k0]))
    sizes = []
    for k in range(1, 10):
        pk, ck = _features(world, traj[k], 200 + k)
        guess = synth.perturb_pose(traj[k], dpos=0.05, drot_deg=0.5, seed=k)
        Tg = guess.copy()
        ok_g = g.Match(_cluster(pk, ck), Tg)
        ok_o, To, st_o = o.match(pk, guess, corner=ck)
        assert ok_g == ok_o and g.last_stats.iterations == st_o.iterations, k
        dt, dr = synth.pose_error(Tg, To)
        assert dt < POS_TOL and dr < ROT_TOL, (k, dt, dr)
        _same_map_size(g, o, k)
        sizes.append(o.map_points)
    assert max(sizes) > min(sizes)
