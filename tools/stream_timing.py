#!/usr/bin/env python
"""Per-scan wall time of a mapping-mode stream (LOAM-iVox): Match + insertion rule + map insert, with the insert path taken."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from funny_lidar_slam_b200 import _abi, synth  # noqa: E402
from funny_lidar_slam_b200.registration import PointcloudCluster, Registration  # noqa: E402

world = synth.make_world()
traj = synth.trajectory(64)
reg = Registration(_abi.default_config(_abi.FLS_P2PLANE_IVOX, localization_mode=0))
first = synth.transform_points(synth.make_scan(world, traj[0], "hdl64", seed=500)["points"], traj[0])
reg.AddCloudToLocalMap([first])
prev = reg.map_info()
for k in range(1, 25):
    scan = synth.make_scan(world, traj[k], "hdl64", seed=500 + k)["points"]
    T = synth.perturb_pose(traj[k], seed=1500 + k, dpos=0.05, drot_deg=0.5)
    t0 = time.perf_counter()
    reg.Match(PointcloudCluster(planar_cloud=scan), T)
    ms = (time.perf_counter() - t0) * 1e3
    mi = reg.map_info()
    path = "incremental" if mi.incremental_inserts > prev.incremental_inserts else "full"
    print(f"scan {k:2d}: {ms:7.2f} ms wall ({reg.last_stats.gpu_ms:6.2f} ms gpu span of the match part, {reg.last_stats.iterations} its) map {mi.n_points:7d} pts "
          f"{mi.n_voxels:6d} vox  +{mi.n_points - prev.n_points:6d} pts  {path}", flush=True)
    prev = mi
