#!/usr/bin/env python
"""Regenerates the golden vectors under tests/golden/ from the CPU oracle on seeded synthetic scenes.

The reference cannot be built or imported here and ships no fixture for this path, so these
goldens pin the ORACLE (regression anchor for the restatement), not the reference: parity stays "unpinned".
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from funny_lidar_slam_b200 import FLS_ICP_P2P, FLS_NDT, FLS_P2PLANE_IVOX, default_config, synth  # noqa: E402
from funny_lidar_slam_b200._abi import FLS_FLAG_ITER_LOG  # noqa: E402
from oracle import pyoracle as orc  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def scene16():
    world = synth.make_world()
    traj = synth.trajectory(16)
    scan = synth.make_scan(world, traj[5], "vlp16", seed=5)["points"]
    mp = synth.make_map_from_scans(world, traj[0:12:2], "vlp16", leaf=0.3)
    return dict(world=world, traj=traj, scan=scan, map=mp, truth=traj[5], guess=synth.perturb_pose(traj[5]),
                guess_small=synth.perturb_pose(traj[5], dpos=0.05, drot_deg=0.5))


def run(method, sc, scan, guess, **kw):
    cfg = default_config(method, flags=FLS_FLAG_ITER_LOG, **kw)
    r = orc.Registration(cfg)
    r.add_cloud(sc["map"])
    ok, T, st = r.match(scan, guess)
    log = r.iter_log()
    return dict(T=T, ok=ok, iters=st.iterations, n_valid=st.n_valid, H0=log[0]["H"], g0=log[0]["g"], n_valid0=log[0]["n_valid"],
                scan_checksum=float(np.sum(scan.astype(np.float64))), map_checksum=float(np.sum(sc["map"].astype(np.float64))))


def main():
    sc = scene16()
    np.savez(os.path.join(OUT, "p2plane_scene16_4000.npz"), **run(FLS_P2PLANE_IVOX, sc, sc["scan"][:4000], sc["guess"]))
    np.savez(os.path.join(OUT, "p2plane_scene16.npz"), **run(FLS_P2PLANE_IVOX, sc, sc["scan"], sc["guess"]))
    np.savez(os.path.join(OUT, "ndt_scene16.npz"), **run(FLS_NDT, sc, sc["scan"], sc["guess_small"]))
    np.savez(os.path.join(OUT, "icp_scene16.npz"), **run(FLS_ICP_P2P, sc, sc["scan"], sc["guess"]))
    proj = synth.make_projected_scan(sc["world"], sc["traj"][7], kind="spin", sensor="vlp16", seed=12)
    ci, pi, _ = orc.extract_features(proj["depth"], proj["col"], len(proj["ordered"]), proj["row_start"], proj["row_end"], 1.0, 0.1)
    np.savez(os.path.join(OUT, "features_vlp16.npz"), corner_idx=ci, planar_idx=pi, n=len(proj["ordered"]),
             depth_checksum=float(np.sum(proj["depth"].astype(np.float64))))
    vg = orc.voxel_grid(sc["scan"], 0.4)
    np.savez(os.path.join(OUT, "voxelgrid_scene16_0p4.npz"), n=len(vg), checksum=float(np.sum(vg.astype(np.float64))), first=vg[:8], last=vg[-8:])
    print("golden vectors written to", OUT)


if __name__ == "__main__":
    main()
