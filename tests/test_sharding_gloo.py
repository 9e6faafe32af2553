"""N > 1 host logic on CPU: world_size-2 gloo processes shard the samples, each evaluates its shard (here with the
oracle standing in for the device), and one all-reduce of the per-root accumulators reproduces the single-process sum."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, out_q):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import fdgraph_b200 as fd
    import graphgen
    from oracle import oracle as O

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    raw, _ = fd.flatten(graphgen.random_dag(77, n_leaves=9, n_inner=50, n_roots=4))
    orc = O.Oracle(raw)
    leaf = graphgen.leaf_values(4, orc.n_leaves, total, signed=True)  # every rank builds the same global sample set
    b, e = fd.shard_range(total, world, rank)
    acc = torch.from_numpy(orc.eval(np.ascontiguousarray(leaf[:, b:e])).sum(axis=1))
    dist.all_reduce(acc)  # the ONE collective of the path: sum of the R accumulators
    full = orc.eval(leaf)
    if rank == 0:
        out_q.put((acc.numpy(), full.sum(axis=1), np.abs(full).sum(axis=1), (b, e)))
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [1000, 1001, 7])
def test_sharded_accumulators_allreduce(total):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    got, want, scale, rng0 = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # 1-GPU vs P-GPU sums differ by reassociation only: |delta| <= k * eps * sum |x_i|
    assert (np.abs(got - want) <= 16 * 2.3e-16 * scale).all()


def test_shard_ranges_tile_the_batch():
    import fdgraph_b200 as fd

    for total in (0, 1, 2, 7, 1000, 1001, 1 << 20):
        for world in (1, 2, 3, 4, 8):
            ranges = [fd.shard_range(total, world, r) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == total
            for (b0, e0), (b1, e1) in zip(ranges, ranges[1:]):
                assert e0 == b1 and b0 <= e0
            assert all(b % 2 == 0 for b, _ in ranges if b < total)  # shards start on a pair boundary
            sizes = [e - b for b, e in ranges]
            assert max(sizes) - min(sizes) <= 3  # one pair, plus the odd sample of the tail
    with pytest.raises(ValueError):
        fd.shard_range(10, 2, 2)
