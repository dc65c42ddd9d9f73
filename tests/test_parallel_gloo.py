"""N > 1 host logic on CPU: world_size-2 gloo process group, scans sharded round-robin, ONE all-gather of the packed
results per batch, merged back in scan order.  The per-scan "matcher" is the CPU oracle here (test infrastructure) —
on the GPU box the same driver logic runs with the CUDA library and NCCL (bench.py)."""
import os
import socket

import numpy as np
import pytest

from funny_lidar_slam_b200 import parallel


def test_shard_indices_cover_every_scan_once():
    for n, w in ((8, 2), (9, 2), (3, 4), (16, 8), (0, 2)):
        seen = sorted(i for r in range(w) for i in parallel.shard_indices(n, r, w))
        assert seen == list(range(n))
    with pytest.raises(ValueError):
        parallel.shard_indices(4, 2, 2)


def test_pack_roundtrip_and_single_process_gather():
    T = np.arange(16, dtype=np.float64).reshape(4, 4)
    v = parallel.pack_result(T, True, 7)
    T2, ok, it = parallel.unpack_result(v)
    assert np.array_equal(T, T2) and ok is True and it == 7
    g = parallel.all_gather_results(v[None, :])  # no process group: identity
    assert g.shape == (1, 1, parallel.RESULT_LEN)
    assert parallel.merge_batch(g, 1)[0][2] == 7


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_scans, out_dir):
    import torch.distributed as dist

    from funny_lidar_slam_b200 import FLS_P2PLANE_IVOX, default_config, synth
    from oracle import pyoracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc.set_num_threads(2)
    world_model = synth.make_world()
    traj = synth.trajectory(8)
    mp = synth.make_map_from_scans(world_model, traj[0:8:3], "vlp16", leaf=0.4)  # replicated map
    reg = orc.Registration(default_config(FLS_P2PLANE_IVOX))
    reg.add_cloud(mp)
    mine = parallel.shard_indices(n_scans, rank, world)
    local = []
    for i in mine:
        scan = synth.make_scan(world_model, traj[i], "vlp16", seed=70 + i)["points"][::4]
        ok, T, st = reg.match(scan, synth.perturb_pose(traj[i], seed=700 + i, dpos=0.1, drot_deg=1.0))
        local.append(parallel.pack_result(T, ok, st.iterations))
    gathered = parallel.all_gather_results(np.stack(local))  # the one collective of the batch
    merged = parallel.merge_batch(gathered, n_scans)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), np.stack([parallel.pack_result(*m) for m in merged]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_batch(tmp_path):
    import torch.multiprocessing as mp

    from funny_lidar_slam_b200 import synth
    n_scans, world = 4, 2
    mp.spawn(_worker, args=(world, _free_port(), n_scans, str(tmp_path)), nprocs=world, join=True)
    r0 = np.load(tmp_path / "rank0.npy")
    r1 = np.load(tmp_path / "rank1.npy")
    assert r0.shape == (n_scans, parallel.RESULT_LEN)
    assert np.array_equal(r0, r1)  # every rank ends up with every pose, in scan order
    traj = synth.trajectory(8)
    for i in range(n_scans):
        T, ok, it = parallel.unpack_result(r0[i])
        assert ok and it >= 1
        assert synth.pose_error(T, traj[i])[0] < 0.05  # scan i really is at slot i


def _async_worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 3
    g = parallel.AsyncResultGather(n, device=None, depth=2)
    got = []
    for step in range(5):
        buf = g.begin_step()  # what the GN kernel writes on the GPU box; here the test fills it
        vals = torch.arange(n * parallel.RESULT_LEN, dtype=torch.float64) + 1000.0 * rank + 100000.0 * step
        buf.copy_(vals)
        g.launch()
        prev = g.take_previous()
        assert (prev is None) == (step == 0)  # one step deferred
        if prev is not None:
            got.append(prev)
    got += g.drain()
    assert len(got) == 5 and g.drain() == []
    np.save(os.path.join(out_dir, f"async{rank}.npy"), np.stack(got))
    dist.barrier()
    dist.destroy_process_group()


def test_async_result_gather_two_ranks(tmp_path):
    """The deferred device-buffer all-gather bench.py runs per step (NCCL there, gloo here): step k's results come out at
    step k+1, every rank sees every rank's buffer, buffers rotate without being overwritten early."""
    import torch.multiprocessing as mp

    world = 2
    mp.spawn(_async_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    a0, a1 = np.load(tmp_path / "async0.npy"), np.load(tmp_path / "async1.npy")
    assert a0.shape == (5, world, 3, parallel.RESULT_LEN) and np.array_equal(a0, a1)
    base = np.arange(3 * parallel.RESULT_LEN, dtype=np.float64).reshape(3, parallel.RESULT_LEN)
    for step in range(5):
        for r in range(world):
            assert np.array_equal(a0[step, r], base + 1000.0 * r + 100000.0 * step)


def test_async_result_gather_single_process_identity():
    import torch
    g = parallel.AsyncResultGather(2)
    b = g.begin_step()
    b.copy_(torch.arange(2 * parallel.RESULT_LEN, dtype=torch.float64))
    g.launch()
    assert g.take_previous() is None
    (out,) = g.drain()
    assert out.shape == (1, 2, parallel.RESULT_LEN) and out[0, 1, 0] == parallel.RESULT_LEN
