"""CPU tests of the N > 1 host-side logic: world_size-2 gloo process group (no GPU)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(ROOT, "osqp-python_b200"))
    import torch
    import torch.distributed as dist
    from b200qp import multigpu
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    m = 11
    bounds = [0, 4, 11]
    lo, hi = bounds[rank], bounds[rank + 1]
    rows = np.full((2, m), -1.0)
    rows[0, lo:hi] = np.arange(lo, hi) + 100 * rank
    rows[1, lo:hi] = -np.arange(lo, hi)
    out = multigpu.gather_slabs(rows, lo, hi, m, world, dist, torch, torch.device("cpu"))
    # timing plumbing of bench.py: max over ranks
    t = torch.tensor([10.0 + rank, 5.0 - rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok_cover = True
    try:
        multigpu.gather_slabs(rows, lo, max(lo, hi - 1), m, world, dist, torch, torch.device("cpu"))
        ok_cover = False
    except RuntimeError:
        pass
    # config-5 sharding (b200qp/batchbench.py): ranges tile [0, batch), counters sum, times max
    from b200qp import batchbench
    lo_b, hi_b = batchbench.shard_range(8191, rank, world)
    dev_ms, wall_ms, tot = batchbench.reduce_counts(3.0 + rank, 7.0 - rank, [hi_b - lo_b, 10 * (rank + 1), 1], dist, torch,
                                                    torch.device("cpu"))
    assert (dev_ms, wall_ms) == (4.0, 7.0) and list(tot) == [8191.0, 30.0, 2.0], (dev_ms, wall_ms, tot)
    q.put((rank, out, t.numpy(), ok_cover))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_slabs_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want0 = np.concatenate([np.arange(0, 4), np.arange(4, 11) + 100])
    for rank, out, t, ok_cover in got:
        np.testing.assert_array_equal(out[0], want0)
        np.testing.assert_array_equal(out[1], -np.arange(11.0))
        np.testing.assert_array_equal(t, [11.0, 5.0])
        assert ok_cover


def test_batch_shard_ranges_tile_the_batch():
    sys.path.insert(0, os.path.join(ROOT, "osqp-python_b200"))
    from b200qp import batchbench
    for batch in (1, 7, 8, 8191, 8192):
        for world in (1, 2, 3, 4, 8):
            edges = [batchbench.shard_range(batch, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == batch
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1


def test_reference_arm_is_rank0_only(monkeypatch, capsys):
    """bench.py --impl reference: ranks other than 0 exit without work."""
    sys.path.insert(0, ROOT)
    import bench
    monkeypatch.setenv("RANK", "1")
    monkeypatch.setenv("WORLD_SIZE", "2")
    bench.run_reference(type("A", (), {"gpus": 2, "steps": 1, "warmup": 0})())
    assert capsys.readouterr().out == ""
