"""Sample sharding for multi-GPU runs (one process per GPU).

The graph evaluator is data-parallel over Monte-Carlo samples and nothing else: the program is
replicated, rank r evaluates its own contiguous range of samples, and the only exchange is ONE sum of the R per-root
accumulators at the end (`fdg_allreduce`, NCCL).  `eval` mode (per-sample roots kept) needs no exchange at all.
"""
from __future__ import annotations

from typing import Tuple


def shard_range(total: int, world: int, rank: int) -> Tuple[int, int]:
    """[begin, end) of the samples rank `rank` evaluates: contiguous, sizes differ by at most one pair, even-sized except
    possibly the last shard (two samples per thread read aligned pairs)."""
    if world < 1 or not (0 <= rank < world) or total < 0:
        raise ValueError("bad shard request")
    pairs = (total + 1) // 2
    base, extra = divmod(pairs, world)
    begin_pair = rank * base + min(rank, extra)
    end_pair = begin_pair + base + (1 if rank < extra else 0)
    return min(2 * begin_pair, total), min(2 * end_pair, total)


def allreduce_accumulators(acc, comm=None, stream: int = 0) -> None:
    """In-place sum over ranks of a float64 CUDA tensor of per-root accumulators through the C ABI (`fdg_allreduce`)."""
    from . import _capi

    if comm is None:
        return
    _capi.check(_capi.lib().fdg_allreduce(comm, acc.data_ptr(), acc.numel(), stream))
