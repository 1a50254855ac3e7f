"""Multi-GPU plumbing: one process per GPU, chains sharded, ONE all-reduce of the accumulators.

Chains (temperature x seed replicas) are independent Markov chains, so the path shards with no
data-path exchange: rank r owns a contiguous block of the global chain range and the
only collective is a sum of the per-temperature accumulator block at the end (NCCL over NVLink on
GPUs; gloo in the CPU tests).  Global chain ids feed the Philox counter, so the reduced integers
are identical for any number of ranks.
"""
from __future__ import annotations

import torch


def shard_range(n: int, rank: int, world: int):
    """Contiguous block [c0, c1) of rank `rank`; sizes differ by at most one."""
    if world <= 1:
        return 0, n
    base, rem = divmod(n, world)
    c0 = rank * base + min(rank, rem)
    return c0, c0 + base + (1 if rank < rem else 0)


def allreduce_sum(t: torch.Tensor, world: int) -> torch.Tensor:
    """In-place SUM over ranks when a process group is up (int64 accumulators -> exact)."""
    if world > 1:
        import torch.distributed as td
        if not td.is_initialized():
            raise RuntimeError("world_size > 1 needs torch.distributed to be initialised")
        td.all_reduce(t, op=td.ReduceOp.SUM)
    return t
