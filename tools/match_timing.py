#!/usr/bin/env python
"""Where does one LoamPointToPlaneIVOX Match spend its time?  Wall clock vs GPU span (first to last op on the handle's
stream) vs the fused kernel alone, for the radix-sort key widths FLS_SORT_KEY_BITS = 24 / 20 / 16.  Warm L2 (no flush)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from funny_lidar_slam_b200 import _abi  # noqa: E402
from funny_lidar_slam_b200.registration import Registration  # noqa: E402

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "p2plane_ivox_64"]
mp, scans, truths, guesses = bench.build_scene(wl, 4, lambda m: None)
reg = Registration(bench.make_cfg(wl, 0, len(mp), flags=_abi.FLS_FLAG_PROFILE))
reg.AddCloudToLocalMap([mp])
d = [torch.from_numpy(s).cuda() for s in scans]
for bits in ("24", "20", "16"):
    os.environ["FLS_SORT_KEY_BITS"] = bits
    for i in range(4):
        reg.match_device(d[i % 4].data_ptr(), d[i % 4].shape[0], guesses[i % 4].copy())
    wall, g, k, its = [], [], [], []
    for i in range(12):
        T = guesses[i % 4].copy()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reg.match_device(d[i % 4].data_ptr(), d[i % 4].shape[0], T)
        t1 = time.perf_counter()
        st = reg.last_stats
        wall.append((t1 - t0) * 1e6)
        g.append(st.gpu_ms * 1e3)
        k.append(st.kernel_ms * 1e3)
        its.append(st.iterations)
    print("key bits", bits, "| wall us", np.round(np.median(wall), 1), "| gpu span us", np.round(np.median(g), 1), "| kernel us",
          np.round(np.median(k), 1), "| kernel/iter", np.round(np.sum(k) / np.sum(its), 2), "| launches", st.gpu_launches)
