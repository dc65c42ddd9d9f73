#!/usr/bin/env python
"""Device-side phase timestamps of the fused LOAM-iVox kernel (FLS_DEBUG_TIMING): one single-scan Match and one batch of 8."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["FLS_DEBUG_TIMING"] = "1"
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from funny_lidar_slam_b200 import _abi  # noqa: E402
from funny_lidar_slam_b200.registration import Registration  # noqa: E402

wl = bench.WORKLOADS["p2plane_ivox_64"]
mp, scans, truths, guesses = bench.build_scene(wl, 8, lambda m: None)
dev = torch.device("cuda", 0)
d = [torch.from_numpy(s).to(dev) for s in scans]
reg = Registration(bench.make_cfg(wl, 0, len(mp), flags=_abi.FLS_FLAG_PROFILE))
reg.AddCloudToLocalMap([mp])
for rep in range(2):
    print(f"--- single scan, call {rep}", file=sys.stderr, flush=True)
    reg.match_device(d[0].data_ptr(), d[0].shape[0], guesses[0].copy())
for rep in range(2):
    print(f"--- batch of 8, call {rep} (scan 0 shown)", file=sys.stderr, flush=True)
    reg.match_batch_device([x.data_ptr() for x in d], [x.shape[0] for x in d], np.stack(guesses))
