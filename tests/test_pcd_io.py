"""PCD files (pcl::io::loadPCDFile / savePCDFileBinary as used by include/common/keyframe.h:24-74 and localization.cpp:283-300
upstream): host-side reader / writer of the product library — no device needed."""
import numpy as np


def test_binary_round_trip(tmp_path):
    from funny_lidar_slam_b200.registration import pcd_read, pcd_write
    rng = np.random.default_rng(0)
    c = rng.normal(0, 30, (5000, 4)).astype(np.float32)
    p = tmp_path / "cloud.pcd"
    pcd_write(p, c)
    head = open(p, "rb").read(400)
    assert b"FIELDS x y z intensity" in head and b"DATA binary" in head and b"POINTS 5000" in head
    assert np.array_equal(pcd_read(p), c)


def test_reads_ascii_and_extra_fields(tmp_path):
    from funny_lidar_slam_b200.registration import pcd_read
    p = tmp_path / "a.pcd"
    rows = [(1.5, -2.25, 3.0, 7, 0.5), (0.125, 4.0, -1.0, 9, 0.25)]
    with open(p, "w") as f:
        f.write("# .PCD v0.7\nVERSION 0.7\nFIELDS x y z ring intensity\nSIZE 4 4 4 2 4\nTYPE F F F U F\nCOUNT 1 1 1 1 1\nWIDTH 2\nHEIGHT 1\n"
                "VIEWPOINT 0 0 0 1 0 0 0\nPOINTS 2\nDATA ascii\n")
        for r in rows:
            f.write(" ".join(str(v) for v in r) + "\n")
    out = pcd_read(p)
    assert np.array_equal(out, np.array([[1.5, -2.25, 3.0, 0.5], [0.125, 4.0, -1.0, 0.25]], np.float32))
    # binary with a field between z and intensity (PointXYZIRT-like records)
    q = tmp_path / "b.pcd"
    rec = np.zeros(3, dtype=[("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("ring", "<u2"), ("intensity", "<f4"), ("time", "<f4")])
    rec["x"], rec["y"], rec["z"], rec["ring"], rec["intensity"], rec["time"] = [1, 2, 3], [4, 5, 6], [7, 8, 9], [1, 2, 3], [0.1, 0.2, 0.3], [0, 1, 2]
    with open(q, "wb") as f:
        f.write(b"VERSION 0.7\nFIELDS x y z ring intensity time\nSIZE 4 4 4 2 4 4\nTYPE F F F U F F\nCOUNT 1 1 1 1 1 1\nWIDTH 3\nHEIGHT 1\nPOINTS 3\nDATA binary\n")
        f.write(rec.tobytes())
    out = pcd_read(q)
    assert np.allclose(out, np.array([[1, 4, 7, 0.1], [2, 5, 8, 0.2], [3, 6, 9, 0.3]], np.float32))


def test_errors(tmp_path):
    import pytest

    from funny_lidar_slam_b200._lib import FlsError
    from funny_lidar_slam_b200.registration import pcd_read
    with pytest.raises(FlsError):
        pcd_read(tmp_path / "missing.pcd")
    p = tmp_path / "c.pcd"
    open(p, "w").write("VERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH 1\nHEIGHT 1\nPOINTS 1\nDATA binary_compressed\n")
    with pytest.raises(FlsError):
        pcd_read(p)
