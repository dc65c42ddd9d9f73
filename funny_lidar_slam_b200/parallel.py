"""Scan-sharded multi-GPU driver logic.

Independent scans against a static, replicated map shard one-per-GPU with NO data-path collective; the only exchange is
ONE all-gather of the per-scan results {4x4 pose, converged, iterations} per batch (NCCL over NVLink for CUDA tensors,
gloo in the CPU tests).  Mapping-mode streams are a sequential chain (scan k's guess and map depend on k-1) and do not
shard: "replicas only".
"""
from __future__ import annotations

import numpy as np

RESULT_LEN = 18  # 16 pose entries (row-major 4x4) + converged + iterations


def shard_indices(n_items: int, rank: int, world: int) -> list[int]:
    """Scan b*world + rank goes to `rank` (the assignment of SURVEY.md §8e): round-robin, no rank idles unless n < world."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return list(range(rank, n_items, world))


def pack_result(T: np.ndarray, converged: bool, iterations: int) -> np.ndarray:
    out = np.empty(RESULT_LEN, np.float64)
    out[:16] = np.asarray(T, np.float64).reshape(-1)
    out[16] = 1.0 if converged else 0.0
    out[17] = float(iterations)
    return out


def unpack_result(v) -> tuple[np.ndarray, bool, int]:
    v = np.asarray(v, np.float64)
    return v[:16].reshape(4, 4).copy(), bool(v[16] > 0.5), int(round(v[17]))


def all_gather_results(local: np.ndarray, device=None):
    """One all-gather of this rank's packed results (shape (k, RESULT_LEN)); returns (world, k, RESULT_LEN).
    Works on whatever backend the default process group uses; without an initialised group it is the identity."""
    import torch
    import torch.distributed as dist

    t = torch.as_tensor(np.ascontiguousarray(local, np.float64))
    if device is not None:
        t = t.to(device)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return t.unsqueeze(0).cpu().numpy()
    bufs = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(bufs, t)
    return torch.stack(bufs).cpu().numpy()


def merge_batch(gathered: np.ndarray, n_items: int) -> list[tuple[np.ndarray, bool, int]]:
    """Undo the round-robin sharding: result of scan i is gathered[i % world][i // world]."""
    world = gathered.shape[0]
    return [unpack_result(gathered[i % world][i // world]) for i in range(n_items)]
