"""GPU parity: K7 VoxelGridCloud on the device vs the oracle's PCL restatement — bit-exact."""
import numpy as np
import pytest

from tests.conftest import to_pcl

pytestmark = pytest.mark.gpu


def _both(points, leaf):
    from funny_lidar_slam_b200.registration import voxel_grid
    from oracle import pyoracle as orc
    return voxel_grid(points, leaf), orc.voxel_grid(points, leaf)


@pytest.mark.parametrize("leaf", [0.2, 0.4, 0.5, 1.0])
def test_scan_bit_exact(scene64, leaf):
    g, o = _both(scene64["scan"], leaf)
    assert g.shape == o.shape and len(g) > 1000
    assert np.array_equal(g.view(np.uint32), o.view(np.uint32))


def test_pcl_layout_and_random_cloud():
    rng = np.random.default_rng(7)
    pts = np.concatenate([rng.uniform(-30, 30, (50000, 3)), rng.uniform(0, 255, (50000, 1))], 1).astype(np.float32)
    from funny_lidar_slam_b200.registration import voxel_grid
    from oracle import pyoracle as orc
    o = orc.voxel_grid(pts, 0.7)
    assert np.array_equal(voxel_grid(pts, 0.7).view(np.uint32), o.view(np.uint32))
    assert np.array_equal(voxel_grid(to_pcl(pts), 0.7).view(np.uint32), o.view(np.uint32))


def test_edge_cases():
    from funny_lidar_slam_b200.registration import voxel_grid
    assert len(voxel_grid(np.zeros((0, 4), np.float32), 0.5)) == 0
    one = np.array([[1.5, -2.25, 3.0, 9.0]], np.float32)
    assert np.array_equal(voxel_grid(one, 0.5), one)
    dup = np.repeat(one, 1000, axis=0)
    g, o = _both(dup, 0.5)
    assert np.array_equal(g.view(np.uint32), o.view(np.uint32)) and len(g) == 1
    # leaf too small for the extent: PCL warns and returns the input unchanged (dx*dy*dz > INT_MAX)
    rng = np.random.default_rng(3)
    wide = np.concatenate([rng.uniform(-400, 400, (2000, 3)), np.ones((2000, 1))], 1).astype(np.float32)
    g, o = _both(wide, 0.05)
    assert len(o) == len(wide) and np.array_equal(g, o)
    # negative coordinates straddling zero, points exactly on cell borders
    grid = np.stack(np.meshgrid(np.arange(-3, 3, 0.25), np.arange(-3, 3, 0.25), [0.0, 0.5], indexing="ij"), -1).reshape(-1, 3)
    pts = np.concatenate([grid, np.arange(len(grid))[:, None]], 1).astype(np.float32)
    g, o = _both(pts, 0.5)
    assert np.array_equal(g.view(np.uint32), o.view(np.uint32))
