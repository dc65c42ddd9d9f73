"""Scan-sharded multi-GPU driver logic.

Independent scans against a static, replicated map shard across GPUs with NO data-path collective; the only exchange is
ONE all-gather of the per-scan results {4x4 pose, converged, iterations} per batch (`ncclAllGather` over NVLink for CUDA
tensors, gloo in the CPU tests).  Mapping-mode streams are a sequential chain (scan k's guess and map depend on k-1) and do
not shard: "replicas only".

The packed result is what the Gauss-Newton kernel itself writes when a scan stops (fls_set_result_buffer_device,
include/fls_b200.h): 16 doubles of Eigen `Mat4d` memory (column-major), converged, iterations.  `AsyncResultGather`
keeps that buffer on the device, issues the all-gather asynchronously on the collective's own stream and hands the
gathered batch out ONE step later, so no rank ever waits for a slower one inside its own step (round 1 staged the poses
through the host and blocked on the collective every step: 0.37 efficiency at 8 GPUs).
"""
from __future__ import annotations

import numpy as np

RESULT_LEN = 18  # 16 pose entries (column-major 4x4, Eigen Mat4d memory) + converged + iterations


def shard_indices(n_items: int, rank: int, world: int) -> list[int]:
    """Scan b*world + rank goes to `rank` (the assignment of SURVEY.md §8e): round-robin, no rank idles unless n < world."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return list(range(rank, n_items, world))


def pack_result(T: np.ndarray, converged: bool, iterations: int) -> np.ndarray:
    out = np.empty(RESULT_LEN, np.float64)
    out[:16] = np.asarray(T, np.float64).T.reshape(-1)  # column-major, like the device-side writer (fls_gn.cuh)
    out[16] = 1.0 if converged else 0.0
    out[17] = float(iterations)
    return out


def unpack_result(v) -> tuple[np.ndarray, bool, int]:
    v = np.asarray(v, np.float64)
    return v[:16].reshape(4, 4).T.copy(), bool(v[16] > 0.5), int(round(v[17]))


def _all_gather(out, inp, async_op: bool):
    """all_gather_into_tensor where the backend has it (NCCL), list form otherwise (gloo)."""
    import torch.distributed as dist

    out, inp = out.view(-1), inp.view(-1)
    try:
        return dist.all_gather_into_tensor(out, inp, async_op=async_op)
    except (RuntimeError, NotImplementedError):
        w = dist.get_world_size()
        return dist.all_gather(list(out.view(w, -1).unbind(0)), inp, async_op=async_op)


def all_gather_results(local: np.ndarray, device=None):
    """One (blocking) all-gather of this rank's packed results (shape (k, RESULT_LEN)); returns (world, k, RESULT_LEN).
    Works on whatever backend the default process group uses; without an initialised group it is the identity."""
    import torch
    import torch.distributed as dist

    t = torch.as_tensor(np.ascontiguousarray(local, np.float64))
    if device is not None:
        t = t.to(device)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return t.unsqueeze(0).cpu().numpy()
    world = dist.get_world_size()
    out = torch.empty((world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
    _all_gather(out, t.contiguous(), False)
    return out.cpu().numpy()


class AsyncResultGather:
    """Device-resident, one-step-deferred all-gather of the per-batch results.

    Usage per step:   buf = g.begin_step()          # device tensor the Match writes into (pass buf.data_ptr() to
                                                     # Registration.set_result_buffer_device once per slot)
                      ... Match (synchronous for the caller: the buffer is complete when it returns) ...
                      g.launch()                     # async all-gather of this step's buffer
                      prev = g.take_previous()       # results of the PREVIOUS step, gathered while this one ran
    `drain()` returns the last outstanding batch.  `depth` buffers rotate, so a buffer is never rewritten while its
    collective may still be reading it.  Without a process group (or world 1) the gather is the identity and costs nothing."""

    def __init__(self, n_scans: int, device=None, depth: int = 2):
        import torch
        import torch.distributed as dist

        self._torch = torch
        self.world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        self.n = int(n_scans)
        self.depth = max(2, int(depth))
        self.local = [torch.zeros(self.n * RESULT_LEN, dtype=torch.float64, device=device) for _ in range(self.depth)]
        self.out = [torch.zeros(self.world * self.n * RESULT_LEN, dtype=torch.float64, device=device) for _ in range(self.depth)]
        self.work = [None] * self.depth
        self.cur = 0       # slot of the step being computed
        self.pending = []  # slots whose collective has been launched and not collected, oldest first

    def begin_step(self):
        s = self.cur
        if self.work[s] is not None:  # the slot's previous collective must have finished before the buffer is rewritten
            self.work[s].wait()
            self.work[s] = None
        return self.local[s]

    def launch(self):
        s = self.cur
        if self.world > 1:
            self.work[s] = _all_gather(self.out[s], self.local[s], True)
        self.pending.append(s)
        self.cur = (s + 1) % self.depth

    def _collect(self, s: int) -> np.ndarray:
        if self.world > 1:
            if self.work[s] is not None:
                self.work[s].wait()
                self.work[s] = None
            src = self.out[s]
        else:
            src = self.local[s]
        return src.cpu().numpy().reshape(self.world, self.n, RESULT_LEN).copy()

    def take_previous(self):
        """Gathered results of the oldest step whose successor has been launched; None until two steps are in flight."""
        if len(self.pending) < 2:
            return None
        return self._collect(self.pending.pop(0))

    def drain(self):
        out = [self._collect(s) for s in self.pending]
        self.pending = []
        return out


def merge_batch(gathered: np.ndarray, n_items: int) -> list[tuple[np.ndarray, bool, int]]:
    """Undo the round-robin sharding: result of scan i is gathered[i % world][i // world]."""
    world = gathered.shape[0]
    return [unpack_result(gathered[i % world][i // world]) for i in range(n_items)]
