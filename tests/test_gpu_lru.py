"""GPU parity: LRU eviction of the incremental maps at capacity (8f-1) — IncrementalNDT::AddCloudToLocalMap
(incremental_ndt.h:193-214) and IVoxMap::AddPoints (ivox_map.cpp:122-143) — against the oracle's sequential lists, voxel sets
compared after every scan of a mapping-mode stream that crosses the capacity."""
import numpy as np
import pytest

from funny_lidar_slam_b200 import FLS_NDT, FLS_P2PLANE_IVOX, default_config, synth

pytestmark = pytest.mark.gpu
POS_TOL, ROT_TOL = 1e-4, 1e-4


def _pair(cfg):
    from funny_lidar_slam_b200.registration import Registration
    from oracle import pyoracle as orc
    return Registration(cfg), orc.Registration(cfg)


def _keyset(keys):
    return set(map(tuple, np.asarray(keys, np.int64).tolist()))


def _ivox_keys(pts, res=0.5):
    v = (pts[:, :3].astype(np.float32) * np.float32(1.0 / res)).astype(np.float32)
    k = np.where(v >= 0, np.floor(v + np.float32(0.5)), np.ceil(v - np.float32(0.5)))  # roundf: half away from zero
    return k.astype(np.int64)


def test_ndt_stream_crosses_capacity(world, traj):
    from funny_lidar_slam_b200.registration import PointcloudCluster
    first = synth.make_scan(world, traj[0], "vlp16", seed=40)["points"]
    first_w = synth.transform_points(first, traj[0])
    probe_cfg = default_config(FLS_NDT, localization_mode=0, max_iterations=15)
    from oracle import pyoracle as orc
    probe = orc.Registration(probe_cfg)
    probe.add_cloud(first_w)
    cap = int(probe.map_voxels * 1.25) + 8  # the second or third scan already overflows it
    cfg = default_config(FLS_NDT, localization_mode=0, max_iterations=15, ndt_capacity=cap)
    g, o = _pair(cfg)
    g.AddCloudToLocalMap([first_w])
    o.add_cloud(first_w)
    assert _keyset(g.voxel_keys()) == _keyset(o.ndt_dump()[0])
    evicted_any = False
    seen = _keyset(o.ndt_dump()[0])
    for k in range(1, 11):
        scan = synth.make_scan(world, traj[k], "vlp16", seed=40 + k)["points"]
        guess = synth.perturb_pose(traj[k], seed=400 + k, dpos=0.05, drot_deg=0.5)
        Tg = guess.copy()
        ok_g = g.Match(PointcloudCluster(ordered_cloud=scan), Tg)
        ok_o, To, st_o = o.match(scan, guess)
        assert ok_g == ok_o, k
        assert g.last_stats.iterations == st_o.iterations, k
        dt, dr = synth.pose_error(Tg, To)
        assert dt < POS_TOL and dr < ROT_TOL, (k, dt, dr)
        ko, kg = _keyset(o.ndt_dump()[0]), _keyset(g.voxel_keys())
        assert g.map_info().n_voxels == o.map_voxels == len(ko), k
        assert kg == ko, (k, len(kg - ko), len(ko - kg))
        assert len(ko) < cap
        evicted_any = evicted_any or bool(seen - ko)
        seen |= ko
    assert evicted_any, "the stream never crossed the capacity: the test is not testing the eviction"


def test_ivox_add_points_crosses_capacity(world, traj):
    """IVoxMap::AddPoints with a small capacity, open loop (the same clouds enter both maps: with pose feedback a one-ulp difference
    of an inserted point near a voxel face changes which voxels exist, and under eviction that avalanches).  After every cloud:
    voxel sets, point multisets and 5-NN answers must be identical."""
    from funny_lidar_slam_b200.registration import Registration
    from oracle import pyoracle as orc
    clouds = [synth.transform_points(synth.make_scan(world, traj[k], "vlp16", seed=60 + k)["points"], traj[k]) for k in range(9)]
    n_first = len(_keyset(_ivox_keys(clouds[0])))
    cap = int(n_first * 1.15) + 8
    g = Registration(default_config(FLS_P2PLANE_IVOX, ivox_capacity=cap))
    o = orc.IVox(0.5, 2, cap)  # NEARBY18
    seen, evicted_any, recreated_any = set(), False, False
    rng = np.random.default_rng(3)
    for k, c in enumerate(clouds):
        # shuffle so that voxels are touched at scattered times of the call: exercises victims that are touched again later
        c = c[rng.permutation(len(c))]
        g.ivox_add_points(c)
        o.add(c)
        mg = g.map_points()
        assert g.map_info().n_voxels == o.num_voxels, k
        assert len(mg) == o.num_points, k
        kg = _keyset(g.voxel_keys())
        assert len(kg) == o.num_voxels and len(kg) < cap
        assert kg == _keyset(_ivox_keys(mg)), k  # the table and the point array agree
        q = c[rng.integers(0, len(c), 2000)] + rng.normal(0, 0.05, (2000, 4)).astype(np.float32)
        pg, ng = g.ivox_knn(q)
        po, no = o.closest(q)
        assert np.array_equal(ng, no), k
        for i in range(2000):
            a, b = pg[i, :ng[i], :3], po[i, :no[i], :3]
            assert np.array_equal(a[np.lexsort(a.T)], b[np.lexsort(b.T)]), (k, i)
        evicted_any = evicted_any or bool(seen - kg)
        recreated_any = recreated_any or (k > 0 and bool((seen - prev) & kg))
        seen |= kg
        prev = kg
    assert evicted_any, "the stream never crossed the capacity: the test is not testing the eviction"


def test_ivox_incremental_insert_equals_full_build(world, traj):
    """Mapping-mode handles insert incrementally (touched voxels + the centres around them are rewritten at the end of the arrays,
    cost independent of the map size).  After every insert the map must answer 5-NN queries exactly like the oracle's IVoxMap and
    like a handle that rebuilds everything (localization-mode handle, same points through fls_ivox_add_points)."""
    from funny_lidar_slam_b200.registration import Registration
    from oracle import pyoracle as orc
    clouds = [synth.transform_points(synth.make_scan(world, traj[k], "vlp16", seed=80 + k)["points"], traj[k]) for k in range(10)]
    g_inc = Registration(default_config(FLS_P2PLANE_IVOX, localization_mode=0))
    g_full = Registration(default_config(FLS_P2PLANE_IVOX))
    o = orc.IVox(0.5, 2, 1000000)
    rng = np.random.default_rng(9)
    g_inc.AddCloudToLocalMap([clouds[0]])  # first cloud: full build
    g_full.ivox_add_points(clouds[0])
    o.add(clouds[0])
    for k in range(1, 10):
        c = clouds[k][rng.permutation(len(clouds[k]))][: 2000 + 1500 * k]  # inserts of growing size, scattered voxels
        g_inc.ivox_add_points(c)
        g_full.ivox_add_points(c)
        o.add(c)
        mi, mf = g_inc.map_info(), g_full.map_info()
        assert mi.n_points == mf.n_points == o.num_points and mi.n_voxels == mf.n_voxels == o.num_voxels, k
        assert _keyset(g_inc.voxel_keys()) == _keyset(g_full.voxel_keys()), k
        q = c[rng.integers(0, len(c), 3000)] + rng.normal(0, 0.2, (3000, 4)).astype(np.float32)
        pi, ni = g_inc.ivox_knn(q)
        pf, nf = g_full.ivox_knn(q)
        po, no = o.closest(q)
        assert np.array_equal(ni, nf) and np.array_equal(ni, no), k
        assert np.array_equal(pi, pf), k  # same candidate order -> identical answers, slot by slot
        for i in range(0, 3000, 5):
            a, b = pi[i, :ni[i], :3], po[i, :no[i], :3]
            assert np.array_equal(a[np.lexsort(a.T)], b[np.lexsort(b.T)]), (k, i)
    mi = g_inc.map_info()
    # (the inserts here are large next to the young map, so its slack runs out a few times; a long stream settles on the incremental path)
    assert mi.incremental_inserts >= 4 and mi.incremental_inserts + mi.full_builds == 10, (mi.incremental_inserts, mi.full_builds)
    assert g_full.map_info().incremental_inserts == 0
