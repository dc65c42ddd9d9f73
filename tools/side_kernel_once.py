#!/usr/bin/env python
"""Runs a few Matches of one side kernel (for ncu): python tools/side_kernel_once.py ndt128|ndt64|icp|kdtree"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from funny_lidar_slam_b200 import _abi, synth  # noqa: E402
from funny_lidar_slam_b200._mem import tune_malloc  # noqa: E402
from funny_lidar_slam_b200.registration import Registration  # noqa: E402

tune_malloc()
which = sys.argv[1] if len(sys.argv) > 1 else "ndt128"
method, sensor, dpos, drot, extra = {
    "ndt64": (_abi.FLS_NDT, "hdl64", 0.05, 0.5, dict(ndt_capacity=2000000)),
    "ndt128": (_abi.FLS_NDT, "os128", 0.05, 0.5, dict(ndt_capacity=2000000, source_cloud_filter_size=0.01, max_iterations=10,
                                                      position_converge_thres=0.0, rotation_converge_thres=0.0)),
    "icp": (_abi.FLS_ICP_P2P, "vlp16", 0.3, 3.0, {}),
    "kdtree": (_abi.FLS_P2PLANE_KNN, "vlp16", 0.1, 1.0, {}),
}[which]
world = synth.make_world()
traj = synth.trajectory(16)
mp = synth.make_surface_map(world, spacing=0.3, seed=4321)
reg = Registration(_abi.default_config(method, **extra))
reg.AddCloudToLocalMap([mp])
for i in range(4):
    sc = synth.make_scan(world, traj[3 + 2 * i], sensor, seed=300 + i)["points"]
    d = torch.from_numpy(sc).cuda()
    T = synth.perturb_pose(traj[3 + 2 * i], seed=900 + i, dpos=dpos, drot_deg=drot)
    ok = reg.match_device(d.data_ptr(), len(sc), T)
    print(which, i, ok, reg.last_stats.iterations, reg.last_stats.n_source, f"{reg.last_stats.gpu_ms:.3f} ms")
