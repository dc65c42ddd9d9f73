"""Host-side mirror of loam::FeatureExtractor (include/loam/feature_extractor.h:15-45 upstream) over the C ABI.

`FeatureExtractor(corner_thr, planar_thr).ExtractFeatures(cluster)` fills cluster.corner_cloud / cluster.planar_cloud
from the projector's arrays exactly as upstream's ExtractFeatures(PointcloudCluster&) does, and also keeps the
index lists (into ordered_cloud) it was built from.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._abi import FlsFeatureCfg, FlsMatchStats
from ._lib import check, lib
from .registration import PointcloudCluster


class FeatureExtractor:
    def __init__(self, corner_thr: float, planar_thr: float, lidar_horizontal_scan: int = 0, lidar_vertical_scan: int = 0, device: int = 0):
        self.cfg = FlsFeatureCfg(float(corner_thr), float(planar_thr), int(device), 0)
        self.last_stats = FlsMatchStats()
        self.corner_idx = np.zeros(0, np.int32)
        self.planar_idx = np.zeros(0, np.int32)

    def extract_indices(self, depth, col, n: int, row_start, row_end):
        depth = np.ascontiguousarray(depth, np.float32)
        col = np.ascontiguousarray(col, np.int32)
        rs = np.ascontiguousarray(row_start, np.int32)
        re = np.ascontiguousarray(row_end, np.int32)
        V = len(rs)
        ci = np.zeros(120 * V + 16, np.int32)
        pi = np.zeros(n + 6 * V + 16, np.int32)
        nc, npl = C.c_size_t(0), C.c_size_t(0)
        st = FlsMatchStats()
        rc = lib().fls_extract_features(C.byref(self.cfg), depth.ctypes.data_as(C.c_void_p), col.ctypes.data_as(C.c_void_p), n,
                                        rs.ctypes.data_as(C.c_void_p), re.ctypes.data_as(C.c_void_p), V, ci.ctypes.data_as(C.c_void_p),
                                        C.byref(nc), pi.ctypes.data_as(C.c_void_p), C.byref(npl), C.byref(st))
        check(rc, "fls_extract_features")
        self.last_stats = st
        self.corner_idx, self.planar_idx = ci[:nc.value].copy(), pi[:npl.value].copy()
        return self.corner_idx, self.planar_idx

    def ExtractFeatures(self, cluster: PointcloudCluster) -> None:
        n = len(cluster.ordered_cloud)
        ci, pi = self.extract_indices(cluster.point_depth_vec, cluster.point_col_index_vec, n, cluster.row_start_index_vec,
                                      cluster.row_end_index_vec)
        cluster.corner_cloud = np.ascontiguousarray(cluster.ordered_cloud[ci])
        cluster.planar_cloud = np.ascontiguousarray(cluster.ordered_cloud[pi])


class PointcloudProjector:
    """Host-side mirror of loam::PointcloudProjector (include/loam/pointcloud_projector.h upstream): `Project(cluster)`
    reads cluster.extra["raw_cloud"] ((n,4) xyzi or (n,8) pcl records, firing order) and cluster.extra["ring"] and
    fills ordered_cloud / point_depth_vec / point_col_index_vec / row_start_index_vec / row_end_index_vec.
    The IMU de-skew step of upstream is not applied (see fls_project in include/fls_b200.h)."""

    def __init__(self, lidar_horizontal_scan: int, lidar_vertical_scan: int, lidar_horizontal_resolution: float, min_distance: float,
                 max_distance: float, device: int = 0):
        self.H, self.V = int(lidar_horizontal_scan), int(lidar_vertical_scan)
        self.h_res, self.min_d, self.max_d = float(lidar_horizontal_resolution), float(min_distance), float(max_distance)
        self.device = int(device)

    def project_arrays(self, raw, ring):
        raw = np.ascontiguousarray(raw, np.float32)
        if raw.ndim != 2 or raw.shape[1] not in (4, 8):
            raise ValueError("raw cloud must be (n,4) packed xyzi or (n,8) pcl records")
        ring = np.ascontiguousarray(ring, np.int32)
        cells = self.V * self.H
        ordered = np.zeros((cells, 4), np.float32)
        depth = np.zeros(cells, np.float32)
        col = np.zeros(cells, np.int32)
        rs = np.zeros(self.V, np.int32)
        re = np.zeros(self.V, np.int32)
        n_out = C.c_size_t(0)
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        rc = lib().fls_project(self.device, vp(raw), vp(ring), C.c_size_t(len(raw)), C.c_size_t(raw.shape[1] * 4), self.V, self.H,
                               C.c_float(self.h_res), C.c_float(self.min_d), C.c_float(self.max_d), vp(ordered), vp(depth), vp(col), vp(rs),
                               vp(re), C.byref(n_out))
        check(rc, "fls_project")
        return dict(ordered=ordered[:n_out.value].copy(), depth=depth, col=col, row_start=rs, row_end=re, n=n_out.value)

    def Project(self, cluster: PointcloudCluster) -> None:
        out = self.project_arrays(cluster.extra["raw_cloud"], cluster.extra["ring"])
        cluster.ordered_cloud = out["ordered"]
        cluster.point_depth_vec = out["depth"]
        cluster.point_col_index_vec = out["col"]
        cluster.row_start_index_vec = out["row_start"]
        cluster.row_end_index_vec = out["row_end"]
