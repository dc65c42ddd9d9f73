"""GPU parity (8f-2): range gate + IMU de-skew + jump span + voxel filter (fls_preprocess) and the projector with its de-skew hook
(fls_project_imu) against the oracle — bit-exact: the device evaluates the quaternion algebra in the oracle's order."""
import numpy as np
import pytest

from tests.test_oracle_deskew import make_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("with_imu", [True, False])
def test_preprocess_bit_exact(with_imu):
    from funny_lidar_slam_b200.features import preprocess
    from oracle import pyoracle as orc
    raw, imu = make_case(n=120000, seed=7)
    imu = imu if with_imu else None
    g_ord, g_pl = preprocess(raw, imu, 2.0, 60.0, 4, 0.5)
    o_ord, o_pl = orc.preprocess(raw, imu, 2.0, 60.0, 4, 0.5)
    assert len(g_ord) == len(o_ord) and len(g_pl) == len(o_pl) and len(g_ord) > 1000
    assert np.array_equal(g_ord, o_ord)
    assert np.array_equal(g_pl, o_pl)


def test_preprocess_ref_time_outside_buffer():
    from funny_lidar_slam_b200.features import preprocess
    raw, imu = make_case(n=3000)
    imu = dict(imu, ref_time_us=int(imu["t_us"][0]) - 5)
    g_ord, g_pl = preprocess(raw, imu, 2.0, 60.0, 4, 0.5)
    assert len(g_ord) == 0 and len(g_pl) == 0


def test_project_imu_bit_exact():
    from funny_lidar_slam_b200.features import PointcloudProjector, project_imu
    from oracle import pyoracle as orc
    raw5, imu = make_case(n=150000, seed=11)
    rng = np.random.default_rng(2)
    ring = rng.integers(-1, 65, len(raw5)).astype(np.int32)  # a few rows out of range
    V, H = 64, 1800
    h_res = float(np.float32(2 * np.pi / H))
    pr = PointcloudProjector(H, V, h_res, 2.0, 60.0)
    g = project_imu(pr, raw5[:, :4], ring, raw5[:, 4], imu)
    o = orc.project_imu(raw5[:, :4], ring, raw5[:, 4], imu, V, H, h_res, 2.0, 60.0)
    assert g["n"] == o["n"] and g["n"] > 10000
    assert np.array_equal(g["ordered"], o["ordered"])
    assert np.array_equal(g["depth"], o["depth"]) and np.array_equal(g["col"], o["col"])
    assert np.array_equal(g["row_start"], o["row_start"]) and np.array_equal(g["row_end"], o["row_end"])
    # and without an IMU buffer it is the plain projector
    g0 = project_imu(pr, raw5[:, :4], ring, raw5[:, 4], None)
    o0 = orc.project(raw5[:, :4], ring, V, H, h_res, 2.0, 60.0)
    assert g0["n"] == o0["n"] and np.array_equal(g0["ordered"], o0["ordered"])
