#!/usr/bin/env python
"""The batch of test_batch_stress_many_small_and_empty_scans (24 scans incl. empty ones, slots shared by three scans each) run
three times with timings — the shape that once hung the v9 kernel (an empty scan never exhausted its ticket counter)."""
import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
from funny_lidar_slam_b200 import FLS_P2PLANE_IVOX, default_config, synth
from funny_lidar_slam_b200.registration import PointcloudCluster, Registration
world = synth.make_world(); traj = synth.trajectory(16)
mp = synth.make_map_from_scans(world, traj[0:12:2], "vlp16", leaf=0.3)
g = Registration(default_config(FLS_P2PLANE_IVOX)); g.AddCloudToLocalMap([mp])
rng = np.random.default_rng(11)
base = [synth.make_scan(world, traj[3 + k], "vlp16", seed=90 + k)["points"] for k in range(4)]
scans, guesses = [], []
for j in range(24):
    k = j % 4
    n = [0, 17, 40, 333, 1500, 5000, len(base[k])][j % 7]
    sel = np.sort(rng.choice(len(base[k]), n, replace=False)) if n else np.zeros(0, np.int64)
    scans.append(np.ascontiguousarray(base[k][sel])); guesses.append(synth.perturb_pose(traj[3 + k], dpos=0.02 + 0.01 * (j % 5), drot_deg=0.3 + 0.2 * (j % 3), seed=200 + j))
for rep in range(3):
    t0 = time.time(); print("batch start", flush=True)
    try:
        conv, Tb = g.match_batch(scans, np.stack(guesses))
        print("batch done", round(time.time() - t0, 3), "s", [s.iterations for s in g.last_batch_stats], flush=True)
    except Exception as e:
        print("batch error", round(time.time() - t0, 3), "s", e, flush=True)
t0 = time.time()
for j in range(24):
    T = guesses[j].copy(); ok = g.Match(PointcloudCluster(planar_cloud=scans[j]), T)
print("singles done", round(time.time() - t0, 3), flush=True)
