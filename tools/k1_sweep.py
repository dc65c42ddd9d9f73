#!/usr/bin/env python
"""A/B harness for the fused LOAM-iVox kernel (K1): builds the bench scene once, then times the kernel (CUDA events of a
FLS_FLAG_PROFILE handle, L2 flushed before every call) for a list of environment variants given on the command line, e.g.
  python tools/k1_sweep.py "" "FLS_K1=8" "FLS_K1_WARPS=18" "FLS_K1_OPTS=1"
Prints per variant: batch-of-8 kernel us (median of N), single-scan kernel us and GPU span."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from funny_lidar_slam_b200 import _abi  # noqa: E402
from funny_lidar_slam_b200.registration import Registration  # noqa: E402

variants = sys.argv[1:] or [""]
N = int(os.environ.get("SWEEP_N", "8"))
wl = bench.WORKLOADS["p2plane_ivox_64"]
mp, scans, truths, guesses = bench.build_scene(wl, 16, lambda m: None)
dev = torch.device("cuda", 0)
d = [torch.from_numpy(s).to(dev) for s in scans]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
base_env = dict(os.environ)
for var in variants:
    for k in list(os.environ):
        if k.startswith("FLS_") and k not in base_env:
            del os.environ[k]
    for kv in filter(None, var.split(",")):
        k, v = kv.split("=")
        os.environ[k] = v
    reg = Registration(bench.make_cfg(wl, 0, len(mp), flags=_abi.FLS_FLAG_PROFILE))
    reg.AddCloudToLocalMap([mp])
    ks, its = [], []
    for i in range(N + 2):
        ids = [(i * 8 + j) % 16 for j in range(8)]
        flush.zero_()
        torch.cuda.synchronize()
        oks, Ts = reg.match_batch_device([d[j].data_ptr() for j in ids], [d[j].shape[0] for j in ids], np.stack([guesses[j] for j in ids]))
        if i >= 2:
            ks.append(sum(x.kernel_ms for x in reg.last_batch_stats) * 1e3)
            its.append(sum(x.iterations for x in reg.last_batch_stats))
    err = max(bench.synth.pose_error(T, truths[j])[0] for T, j in zip(Ts, ids))
    sk, sg = [], []
    for i in range(6):
        flush.zero_()
        torch.cuda.synchronize()
        T = guesses[i % 16].copy()
        reg.match_device(d[i % 16].data_ptr(), d[i % 16].shape[0], T)
        if i >= 2:
            sk.append(reg.last_stats.kernel_ms * 1e3)
            sg.append(reg.last_stats.gpu_ms * 1e3)
    print(f"{var or 'default':32s} batch8 kernel us: median {np.median(ks):7.1f} min {np.min(ks):7.1f} (iters {np.mean(its):.1f}) | single: kernel {np.median(sk):6.1f} span {np.median(sg):6.1f} | max pos err {err:.4f}", flush=True)
    del reg
