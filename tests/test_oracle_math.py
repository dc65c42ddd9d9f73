"""Oracle pinning (CPU): SO(3)/SE(3) helpers against the ONLY known-answer values the reference ships
(test/math_function_ut.cpp, test/lidar_model_ut.cpp upstream) and the small dense solvers against numpy."""
import numpy as np
import pytest

from oracle import pyoracle as orc


def test_so3_exp_reference_kats():
    # test/math_function_ut.cpp:68-127 upstream
    I = np.eye(3)
    assert np.allclose(orc.so3_exp([0, 0, 0]), I)
    e1 = np.array([1.0, 0, 0])
    hat = np.array([[0, 0, 0], [0, 0, -1], [0, 1, 0]], float)
    assert np.allclose(orc.so3_exp(e1 * np.pi / 2), np.outer(e1, e1) + hat, atol=1e-15)
    for th in (np.pi / 4, 3 * np.pi, -3 * np.pi):
        want = np.cos(th) * I + (1 - np.cos(th)) * np.outer(e1, e1) + np.sin(th) * hat
        assert np.allclose(orc.so3_exp(e1 * th), want, atol=1e-14)
    assert np.allclose(orc.so3_exp(e1).T, orc.so3_exp(-e1), atol=1e-15)


def test_se3_exp_golden_matrix():
    # test/math_function_ut.cpp:135-148 upstream: SE3Exp([0.5,1,1.5 | 0.1,0.001,0.00124])
    T_true = np.array([[0.999998732257249, -0.0011879755058542, 0.00106028208140807, 0.500177323147329],
                       [0.00128789217916111, 0.99500339817527, -0.0998327549124039, 0.923714786259781],
                       [-0.000936385406507498, 0.0998339938791529, 0.995003666751288, 1.54722007984458],
                       [0, 0, 0, 1]])
    T = orc.se3_exp([0.5, 1.0, 1.5, 0.1, 0.001, 0.00124])
    assert np.allclose(T, T_true, rtol=1e-12, atol=1e-14)
    # the rotation block is exactly what the GN update applies
    assert np.allclose(orc.so3_exp([0.1, 0.001, 0.00124]), T_true[:3, :3], rtol=1e-12, atol=1e-14)


def test_rotation_matrix_to_rpy_kat():
    # test/math_function_ut.cpp:160-192 upstream: R = Rz(pi/3) Ry(pi/4) Rx(pi/6)
    def rx(t): return np.array([[1, 0, 0], [0, np.cos(t), -np.sin(t)], [0, np.sin(t), np.cos(t)]])
    def ry(t): return np.array([[np.cos(t), 0, np.sin(t)], [0, 1, 0], [-np.sin(t), 0, np.cos(t)]])
    def rz(t): return np.array([[np.cos(t), -np.sin(t), 0], [np.sin(t), np.cos(t), 0], [0, 0, 1]])
    rpy = orc.rot_to_rpy(rz(np.pi / 3) @ ry(np.pi / 4) @ rx(np.pi / 6))
    assert np.allclose(rpy, [np.pi / 6, np.pi / 4, np.pi / 3], atol=1e-15)


def test_fast_atan2_accuracy_and_projector_column_kat():
    # FastAtan2 "can guarantee angular accuracy of two decimal places" (math_function.h:152-158 upstream)
    rng = np.random.default_rng(0)
    xy = rng.normal(size=(2000, 2)).astype(np.float32)
    got = np.array([orc.fast_atan2f(y, x) for x, y in xy])
    assert np.max(np.abs(np.angle(np.exp(1j * (got - np.arctan2(xy[:, 1], xy[:, 0])))))) < 2e-3
    # column index = round(atan2/h_res) + H/2 for a LeiShen_16-like sensor (test/lidar_model_ut.cpp:9-51 upstream: H=2000, 0.18 deg)
    h_res = np.float32(np.deg2rad(0.18))
    for ang_deg, col in ((0.0, 1000), (90.0, 1500), (-90.0, 500), (179.9, 1999)):
        a = np.deg2rad(ang_deg)
        c = int(np.round(orc.fast_atan2f(np.float32(np.sin(a) * 10), np.float32(np.cos(a) * 10)) / h_res)) + 1000
        assert abs(c - col) <= 1


def test_plane_lstsq_matches_numpy():
    rng = np.random.default_rng(1)
    for _ in range(200):
        n = rng.normal(size=3)
        n /= np.linalg.norm(n)
        d = rng.uniform(0.5, 80)
        base = rng.normal(size=(5, 3)) * 0.3
        P = base - np.outer(base @ n, n) + n * d + rng.normal(size=(5, 3)) * 0.01  # noisy plane n.x = d
        P32 = P.astype(np.float32).astype(np.float64)
        c = orc.lstsq53(P32, -np.ones(5))
        want = np.linalg.lstsq(P32, -np.ones(5), rcond=None)[0]
        assert np.allclose(c, want, rtol=1e-9, atol=1e-12)
    # rank-deficient input (collinear points): basic solution, finite
    P = np.outer(np.arange(1, 6), [1.0, 2.0, 3.0])
    c = orc.lstsq53(P, -np.ones(5))
    assert np.all(np.isfinite(c)) and np.allclose(P @ c, np.linalg.lstsq(P, -np.ones(5), rcond=None)[0] @ P.T, atol=1e-9)


def test_6x6_solvers_match_numpy():
    rng = np.random.default_rng(2)
    for _ in range(50):
        J = rng.normal(size=(40, 6))
        H = J.T @ J
        g = rng.normal(size=6)
        want = np.linalg.solve(H, g)
        assert np.allclose(orc.solve6_fullpiv(H, g), want, rtol=1e-9, atol=1e-12)
        x, det = orc.solve6_lu(H, g)
        assert np.allclose(x, want, rtol=1e-9, atol=1e-12)
        assert np.isclose(det, np.linalg.det(H), rtol=1e-9)
    # singular: LU reports det == 0 (IcpOptimized `continue`), full-pivot returns the basic solution
    x, det = orc.solve6_lu(np.zeros((6, 6)), np.ones(6))
    assert det == 0.0 and np.all(x == 0)
    H = np.diag([1.0, 2.0, 3.0, 0.0, 0.0, 0.0])
    x = orc.solve6_fullpiv(H, np.array([1.0, 2.0, 3.0, 0.0, 0.0, 0.0]))
    assert np.allclose(x, [1, 1, 1, 0, 0, 0])


def test_sym_eig3_matches_numpy():
    rng = np.random.default_rng(3)
    for _ in range(50):
        A = rng.normal(size=(3, 3))
        S = A @ A.T
        lam, V = orc.sym_eig3(S)
        assert np.allclose(lam, np.sort(np.linalg.eigvalsh(S))[::-1], rtol=1e-10)
        assert np.allclose(V @ np.diag(lam) @ V.T, S, atol=1e-10)


@pytest.mark.parametrize("n", [1, 2, 8])
def test_oracle_thread_count_does_not_change_results(n, scene16):
    from funny_lidar_slam_b200 import FLS_P2PLANE_IVOX, default_config
    prev = orc.num_threads()
    try:
        orc.set_num_threads(n)
        r = orc.Registration(default_config(FLS_P2PLANE_IVOX))
        r.add_cloud(scene16["map"])
        ok, T, st = r.match(scene16["scan"][:4000], scene16["guess"])
    finally:
        orc.set_num_threads(prev)
    ref = np.load(__file__.replace("test_oracle_math.py", "golden/p2plane_scene16_4000.npz"))
    assert np.allclose(T, ref["T"], atol=1e-12) and st.iterations == int(ref["iters"])
