"""Oracle pinning (CPU): the kd-tree LOAM plug-ins (LoamPointToPlaneKdtree, LoamFull) and the projector.

Independent checks, none of which shares code with the oracle:
  * one PlanerMatch + SumCoefficient pass restated in numpy / scipy (cKDTree 5-NN, lstsq plane, the gates, J J^T sums);
  * the point-to-plane and point-to-line Jacobians as finite differences of the residual they linearise
    (doc/loam_formula.md:27-157 upstream is the derivation; the perturbation is the LEFT one, R <- Exp(dθ) R, t <- t + dt);
  * golden vectors (tests/golden/make_golden.py) as regression anchors;
  * projector semantics on hand-made inputs (first hit wins, column wrap, row bounds)."""
import os

import numpy as np
import pytest
from scipy.spatial import cKDTree

from funny_lidar_slam_b200 import default_config, synth
from funny_lidar_slam_b200._abi import FLS_FLAG_ITER_LOG, FLS_LOAM_FULL, FLS_P2PLANE_KNN
from oracle import pyoracle as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _features(world, pose, seed):
    proj = synth.make_projected_scan(world, pose, kind="spin", sensor="vlp16", seed=seed)
    n = len(proj["ordered"])
    ci, pi, _ = orc.extract_features(proj["depth"], proj["col"], n, proj["row_start"], proj["row_end"], 1.0, 0.1)
    return proj["ordered"][pi].copy(), proj["ordered"][ci].copy()


def _to_world(pts, T):
    out = pts.copy()
    out[:, :3] = (pts[:, :3].astype(np.float64) @ T[:3, :3].T + T[:3, 3]).astype(np.float32)
    return out


@pytest.fixture(scope="module")
def feature_scene(world, traj):
    maps_p, maps_c = [], []
    for k in (3, 4, 6, 7):
        p, c = _features(world, traj[k], k)
        maps_p.append(_to_world(p, traj[k]))
        maps_c.append(_to_world(c, traj[k]))
    p5, c5 = _features(world, traj[5], 55)
    return dict(maps_p=maps_p, maps_c=maps_c, planar=p5, corner=c5, truth=traj[5], guess=synth.perturb_pose(traj[5], dpos=0.1, drot_deg=1.0))


def _hat(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])


def test_kdtree_first_iteration_against_numpy(feature_scene):
    """PlanerMatch + SumCoefficient of loam_point_to_plane_kdtree.h:204-303, restated with scipy / numpy."""
    cfg = default_config(FLS_P2PLANE_KNN, flags=FLS_FLAG_ITER_LOG, max_iterations=1)
    mp = np.concatenate(feature_scene["maps_p"])
    r = orc.Registration(cfg)
    r.add_cloud(mp)
    src = feature_scene["planar"][::5]
    T = feature_scene["guess"]
    r.match(src, T)
    lg = r.iter_log()[0]
    vmap = orc.voxel_grid(mp, cfg.map_cloud_filter_size)  # what upstream builds its kd-tree on (:78-79)
    tree = cKDTree(vmap[:, :3].astype(np.float64))
    R, t = T[:3, :3], T[:3, 3]
    q = (src[:, :3].astype(np.float64) @ R.T + t).astype(np.float32).astype(np.float64)  # pcl::transformPoint: fp64 math, fp32 store
    _, idx = tree.query(q, k=5)
    H, g, nv, res = np.zeros((6, 6)), np.zeros(6), 0, 0.0
    for i in range(len(src)):
        A = vmap[idx[i], :3].astype(np.float64)
        c = np.linalg.lstsq(A, -np.ones(5), rcond=None)[0]
        cn = np.linalg.norm(c)
        if np.any(np.abs(A @ c + 1.0) / cn > cfg.point_to_planar_thres):
            continue
        n = c / cn
        ps = src[i, :3].astype(np.float64)
        d = (q[i] - A[0]) @ n
        if np.linalg.norm(ps) < 81 * d * d:
            continue
        s = 1.0 if d > 0 else -1.0
        J = np.concatenate([-_hat(R @ ps).T @ n * s, n * s])
        H += np.outer(J, J)
        g += -J * abs(d)
        nv += 1
        res += abs(d)
    assert nv == lg["n_valid"] and nv > 1000
    assert np.allclose(H, lg["H"], rtol=1e-9, atol=1e-9 * np.abs(H).max())
    assert np.allclose(g, lg["g"], rtol=1e-9, atol=1e-9)
    assert res == pytest.approx(lg["sum_residual"], rel=1e-10)


def _left_perturb(T, d):
    th = np.linalg.norm(d[:3])
    K = _hat(d[:3] / th) if th > 0 else np.zeros((3, 3))
    E = np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K
    out = T.copy()
    out[:3, :3] = E @ T[:3, :3]
    out[:3, 3] = T[:3, 3] + d[3:]
    return out


def _fd_gradient(f, eps=1e-6):
    g = np.zeros(6)
    for k in range(6):
        d = np.zeros(6)
        d[k] = eps
        g[k] = (f(d) - f(-d)) / (2 * eps)
    return g


def test_planar_jacobian_is_the_derivative_of_the_plane_distance():
    """One planar point against a 5-point plane patch: H = J J^T, g = -J |d| (loam_point_to_plane_kdtree.h:270-282), so
    J = -g / |d|, and J must be the gradient of |d| under the left perturbation the solver applies (:108-112)."""
    rng = np.random.default_rng(4)
    patch = np.array([[10.0, 2.0, 1.0], [10.4, 2.1, 1.02], [10.1, 2.6, 0.99], [9.7, 1.8, 1.01], [10.3, 1.7, 0.98]])
    filler = rng.uniform(40, 60, (20, 3))  # far away: never among the 5 nearest
    mp = np.zeros((25, 4), np.float32)
    mp[:5, :3] = patch
    mp[5:, :3] = filler
    T = synth.perturb_pose(np.eye(4), dpos=0.5, drot_deg=20.0, seed=3)
    pw = np.array([10.1, 2.1, 1.3])                      # where the source point lands in the map frame
    ps = (T[:3, :3].T @ (pw - T[:3, 3])).astype(np.float32)
    src = np.array([[ps[0], ps[1], ps[2], 0.0]], np.float32)
    cfg = default_config(FLS_P2PLANE_KNN, flags=FLS_FLAG_ITER_LOG, max_iterations=1, map_cloud_filter_size=0.05, point_to_planar_thres=0.2)
    r = orc.Registration(cfg)
    r.add_cloud(mp)
    r.match(src, T)
    lg = r.iter_log()[0]
    assert lg["n_valid"] == 1
    ad = lg["sum_residual"]
    J = -lg["g"] / ad
    vm = orc.voxel_grid(mp, 0.05)[:, :3].astype(np.float64)
    near = vm[np.argsort(np.linalg.norm(vm - pw, axis=1))[:5]]
    c = np.linalg.lstsq(near, -np.ones(5), rcond=None)[0]
    n = c / np.linalg.norm(c)

    def absdist(d):
        Tp = _left_perturb(T, d)
        return abs((Tp[:3, :3] @ ps.astype(np.float64) + Tp[:3, 3] - near[0]) @ n)

    assert ad == pytest.approx(absdist(np.zeros(6)), rel=1e-5)
    assert np.allclose(J, _fd_gradient(absdist), atol=1e-5)
    assert np.allclose(lg["H"], np.outer(J, J), atol=1e-12)


def test_corner_jacobian_is_the_derivative_of_the_line_distance():
    """One corner point against 5 nearly collinear map points (loam_full_kdtree.h:211-273): J = -g / d must be the gradient
    of the point-to-line distance ||(q - c) x n|| with n the principal axis of the neighbours' covariance."""
    rng = np.random.default_rng(5)
    axis = np.array([0.2, 0.1, 1.0])
    axis /= np.linalg.norm(axis)
    line = np.array([12.0, -3.0, 0.5]) + np.outer(np.array([-0.4, -0.2, 0.0, 0.2, 0.4]), axis) + rng.normal(0, 0.004, (5, 3))
    mc = np.zeros((25, 4), np.float32)
    mc[:5, :3] = line
    mc[5:, :3] = rng.uniform(40, 60, (20, 3))
    mpl = np.zeros((30, 4), np.float32)
    mpl[:, :3] = rng.uniform(70, 90, (30, 3))           # planar map far away: contributes nothing
    T = synth.perturb_pose(np.eye(4), dpos=0.5, drot_deg=15.0, seed=9)
    pw = np.array([12.15, -2.9, 0.55])
    ps = (T[:3, :3].T @ (pw - T[:3, 3])).astype(np.float32)
    corner = np.array([[ps[0], ps[1], ps[2], 0.0]], np.float32)
    planar = np.array([[1.0, 1.0, 1.0, 0.0]], np.float32)  # lands nowhere near the planar map: gated by point_search_thres
    cfg = default_config(FLS_LOAM_FULL, flags=FLS_FLAG_ITER_LOG, max_iterations=1)
    r = orc.Registration(cfg)
    r.add_cloud(mpl, mc)
    ok, _, _ = r.match(planar, T, corner=corner)
    assert not ok  # fewer than 50 valid planar points (:174-176) — the corner term is summed all the same
    lg = r.iter_log()[0]
    d0 = lg["sum_residual"]
    J = -lg["g"] / d0
    P = mc[:5, :3].astype(np.float64)
    cen = P.mean(0)
    w, V = np.linalg.eigh((P - cen).T @ (P - cen) / 5.0)
    n = V[:, 2]
    assert w[2] > cfg.line_ratio_thres * w[1]

    def dist(d):
        Tp = _left_perturb(T, d)
        return np.linalg.norm(np.cross(Tp[:3, :3] @ ps.astype(np.float64) + Tp[:3, 3] - cen, n))

    assert d0 == pytest.approx(dist(np.zeros(6)), rel=1e-5)
    assert np.allclose(J, _fd_gradient(dist), atol=1e-5)
    assert np.allclose(lg["H"], np.outer(J, J), atol=1e-12)


@pytest.mark.parametrize("name", ["kdtree_features", "loamfull_features"])
def test_against_golden(feature_scene, name):
    ref = np.load(os.path.join(GOLD, name + ".npz"))
    assert float(np.sum(feature_scene["planar"].astype(np.float64))) == float(ref["planar_checksum"])
    if name == "kdtree_features":
        r = orc.Registration(default_config(FLS_P2PLANE_KNN, flags=FLS_FLAG_ITER_LOG))
        r.add_cloud(np.concatenate(feature_scene["maps_p"]))
        ok, T, st = r.match(feature_scene["planar"], feature_scene["guess"])
    else:
        r = orc.Registration(default_config(FLS_LOAM_FULL, localization_mode=0, flags=FLS_FLAG_ITER_LOG))
        for mp, mc in zip(feature_scene["maps_p"], feature_scene["maps_c"]):
            r.add_cloud(mp, mc)
        ok, T, st = r.match(feature_scene["planar"], feature_scene["guess"], corner=feature_scene["corner"])
    lg = r.iter_log()
    assert ok == bool(ref["ok"]) and st.iterations == int(ref["iters"]) and st.n_valid == int(ref["n_valid"])
    assert np.allclose(lg[0]["H"], ref["H0"], rtol=1e-12, atol=1e-9) and np.allclose(lg[0]["g"], ref["g0"], rtol=1e-12, atol=1e-9)
    assert np.allclose(T, ref["T"], atol=1e-10)
    assert synth.pose_error(T, feature_scene["truth"])[0] < 0.02


def test_loam_full_map_rules():
    """Windows slide independently and the voxel filters only run beyond 5 clouds (loam_full_kdtree.h:75-99)."""
    rng = np.random.default_rng(2)
    cfg = default_config(FLS_LOAM_FULL, localization_mode=0, local_map_size=7, corner_local_map_size=3, map_cloud_filter_size=0.5,
                         corner_map_filter_size=0.5)
    r = orc.Registration(cfg)
    sizes = []
    for k in range(8):
        pl = np.zeros((200, 4), np.float32)
        pl[:, :3] = rng.uniform(0, 3, (200, 3))  # dense: the filter visibly thins the union once it runs
        co = np.zeros((10, 4), np.float32)
        co[:, :3] = rng.uniform(0, 3, (10, 3))
        r.add_cloud(pl, co)
        sizes.append((len(r.map_copy(0)), len(r.map_copy(1))))
    assert [s[0] for s in sizes[:5]] == [200, 400, 600, 800, 1000]   # unfiltered up to 5 clouds
    assert sizes[5][0] < 1200 and sizes[7][0] <= 216                   # filtered: at most 6^3 cells of 0.5 m
    assert [s[1] for s in sizes] == [10, 20, 30, 30, 30, 30, 30, 30]   # corner window of 3 never reaches the filter


def test_projector_semantics():
    V, H = 2, 8
    h_res = float(np.float32(2 * np.pi / H))

    def pt(az, rng_, z=0.0):
        return [rng_ * np.cos(az), rng_ * np.sin(az), z, 7.0]

    raw = np.array([pt(0.0, 10), pt(0.01, 11), pt(np.pi / 4, 5), pt(-np.pi + 1e-3, 6), pt(np.pi - 1e-3, 6.5), pt(0.0, 200.0), pt(0.0, 0.5)],
                   np.float32)
    ring = np.array([0, 0, 1, 1, 1, 0, 0], np.int32)
    o = orc.project(raw, ring, V, H, h_res, 1.0, 100.0)
    # ring 0: both az~0 points fall into column H/2; the first one keeps the cell; out-of-range points are dropped
    # ring 1: az=pi/4 -> column H/2+1; az=-pi -> column 0; az=+pi rounds to H/2+H/2 = H and wraps to 0, already taken
    assert o["n"] == 3
    assert np.array_equal(o["col"][:3], [H // 2, 0, H // 2 + 1])
    assert np.allclose(o["depth"][:3], [10.0, 6.0, 5.0], atol=1e-5)
    assert np.array_equal(o["ordered"][:, 3], [7.0, 7.0, 7.0])
    assert list(o["row_start"]) == [5, 6] and list(o["row_end"]) == [-5, -3]  # count+5 / count-6 per row (:115,:131)
