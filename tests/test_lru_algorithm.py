"""The batch form of upstream's LRU insert (IVoxMap::AddPoints, ivox_map.cpp:122-143; IncrementalNDT::AddCloudToLocalMap,
incremental_ndt.h:193-214) used by the device maps: per call, candidates = existing voxels by ascending last-touch stamp with the
index of the first point of the call that touches them, creation times of the new voxels, and a merge that reproduces the
sequential policy (funny_lidar_slam_b200/csrc/fls_map.cu: lru_simulate).  Checked here against a literal sequential restatement on
random streams — the Python below mirrors the C++ line by line."""
import heapq
from collections import OrderedDict

import numpy as np


def sequential(calls, capacity):
    """Literal upstream policy; returns the voxel -> points map after every call."""
    lru = OrderedDict()  # front = last
    out = []
    for pts in calls:
        for t, key in enumerate(pts):
            if key not in lru:
                lru[key] = []
                lru.move_to_end(key)
                lru[key].append((len(out), t))
                if len(lru) >= capacity:
                    lru.popitem(last=False)
            else:
                lru[key].append((len(out), t))
                lru.move_to_end(key)
        out.append({k: list(v) for k, v in lru.items()})
    return out


def lru_simulate(size0, capacity, cand_first_touch, create_times):
    events = list(create_times)
    heapq.heapify(events)
    victims, recreated = [], []
    size, ci = size0, 0
    while events:
        t = heapq.heappop(events)
        size += 1
        if size < capacity:
            continue
        while ci < len(cand_first_touch) and cand_first_touch[ci] < t:
            ci += 1
        if ci >= len(cand_first_touch):
            return None
        ft = cand_first_touch[ci]
        victims.append(ci)
        recreated.append(ft != 0xFFFFFFFF)
        if ft != 0xFFFFFFFF:
            heapq.heappush(events, ft)
        ci += 1
        size -= 1
    return victims, recreated


def batched(calls, capacity):
    """The device algorithm: state = per voxel its points with (call, index) stamps."""
    vox = {}
    out = []
    for c, pts in enumerate(calls):
        first, last = {}, {}
        for t, key in enumerate(pts):
            first.setdefault(key, t)
            last[key] = t
        existing = sorted(vox, key=lambda k: vox[k][-1])  # ascending stamp of the last point
        creations = [first[k] for k in first if k not in vox]
        if len(vox) + len(creations) >= capacity:
            cand = [first.get(k, 0xFFFFFFFF) for k in existing]
            res = lru_simulate(len(vox), capacity, cand, creations)
            assert res is not None
            for pos in res[0]:
                del vox[existing[pos]]  # the victim's old points go; touched later in the call -> re-created below
        for t, key in enumerate(pts):
            vox.setdefault(key, []).append((c, t))
        out.append({k: list(v) for k, v in vox.items()})
    return out


def test_batched_lru_equals_sequential_on_random_streams():
    rng = np.random.default_rng(5)
    for trial in range(200):
        capacity = int(rng.integers(12, 40))
        n_calls = int(rng.integers(3, 9))
        universe = int(rng.integers(capacity, 4 * capacity))
        calls = []
        for c in range(n_calls):
            centre = rng.integers(0, universe)
            n = int(rng.integers(1, capacity - 2))  # a single call never holds more distinct voxels than the capacity allows
            keys = (centre + rng.integers(-capacity // 3, capacity // 3 + 1, size=3 * n)) % universe
            keys = [int(k) for k in keys]
            # keep the number of distinct voxels of a call below the capacity (upstream is undefined beyond that)
            seen = []
            pts = []
            for k in keys:
                if k not in seen:
                    if len(seen) >= capacity - 2:
                        continue
                    seen.append(k)
                pts.append(k)
            calls.append(pts)
        a = sequential(calls, capacity)
        b = batched(calls, capacity)
        for c, (x, y) in enumerate(zip(a, b)):
            assert set(x) == set(y), (trial, c)
            for k in x:
                assert x[k] == y[k], (trial, c, k)
