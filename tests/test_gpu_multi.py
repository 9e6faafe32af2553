"""N > 1 on real GPUs: the C entry points of the one collective of the path (fdg_comm_unique_id / fdg_comm_init /
fdg_allreduce / fdg_comm_destroy) called by two processes, one per GPU, each with its own shard of one counter-based
sample stream (sharding.shard_range).  The all-reduced accumulators must equal the single-process sums over the union of
the shards to reassociation.  Skipped when the box has a single GPU."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, idfile, total, out_q):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import ctypes
    import time

    import torch

    import fdgraph_b200 as fd
    import graphgen
    from fdgraph_b200 import _capi

    torch.cuda.set_device(rank)
    L = _capi.lib()
    if rank == 0:
        buf = (ctypes.c_ubyte * 128)()
        _capi.check(L.fdg_comm_unique_id(buf))
        with open(idfile + ".tmp", "wb") as fh:
            fh.write(bytes(buf))
        os.replace(idfile + ".tmp", idfile)
    t0 = time.time()
    while not os.path.exists(idfile):
        time.sleep(0.05)
        assert time.time() - t0 < 60
    ident = (ctypes.c_ubyte * 128).from_buffer_copy(open(idfile, "rb").read())
    comm = ctypes.c_void_p()
    _capi.check(L.fdg_comm_init(ctypes.byref(comm), world, rank, ident))
    raw, _ = fd.flatten(graphgen.random_dag(77, n_leaves=9, n_inner=50, n_roots=4))
    ev = fd.compile_raw(raw)
    leaf = graphgen.leaf_values(4, ev.n_leaves, total, signed=True)  # every rank builds the same global sample set ...
    b, e = fd.shard_range(total, world, rank)                         # ... and evaluates its own shard of it
    mine = torch.from_numpy(np.ascontiguousarray(leaf[:, b:e])).cuda()
    acc = torch.zeros(ev.n_roots, dtype=torch.float64, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    ev.accumulate_device(mine.data_ptr(), e - b, e - b, acc.data_ptr(), s)
    _capi.check(L.fdg_allreduce(comm, acc.data_ptr(), ev.n_roots, s))
    torch.cuda.synchronize()
    # the single-GPU answer over the whole set, computed by this rank too
    full = torch.from_numpy(leaf).cuda()
    one = torch.zeros(ev.n_roots, dtype=torch.float64, device="cuda")
    root = torch.zeros(ev.n_roots, total, dtype=torch.float64, device="cuda")
    ev.accumulate_device(full.data_ptr(), total, total, one.data_ptr(), s)
    ev.eval_device(full.data_ptr(), total, root.data_ptr(), total, total, s)
    torch.cuda.synchronize()
    out_q.put((rank, acc.cpu().numpy(), one.cpu().numpy(), root.abs().sum(dim=1).cpu().numpy()))
    _capi.check(L.fdg_comm_destroy(comm))


@pytest.mark.parametrize("total", [100_000, 100_001])
def test_allreduce_of_sharded_accumulators_on_two_gpus(total, tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp

    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    idfile = str(tmp_path / "nccl_id")
    procs = [ctx.Process(target=_worker, args=(r, world, idfile, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, got, one, scale in results:
        # 1-GPU vs 2-GPU sums differ by reassociation only: |delta| <= k eps sum |x_i|
        assert (np.abs(got - one) <= 64 * 2.3e-16 * scale).all(), rank
    assert np.array_equal(results[0][1], results[1][1])  # both ranks hold the same reduced values
