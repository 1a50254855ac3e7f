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


def allgather_rows(t: torch.Tensor, n_total: int, rank: int, world: int) -> torch.Tensor:
    """Rows of every rank's contiguous shard (shard_range) on every rank, in global order: [n_total, ...].
    One all_gather of equally sized (padded) blocks -- the optional per-chain gather of SURVEY 8e, used for
    error bars over replicas on rank 0."""
    if world <= 1:
        return t
    import torch.distributed as td
    if not td.is_initialized():
        raise RuntimeError("world_size > 1 needs torch.distributed to be initialised")
    width = -(-n_total // world)
    pad = torch.zeros((width,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[:t.shape[0]] = t
    parts = [torch.empty_like(pad) for _ in range(world)]
    td.all_gather(parts, pad)
    rows = []
    for r in range(world):
        c0, c1 = shard_range(n_total, r, world)
        rows.append(parts[r][:c1 - c0])
    return torch.cat(rows, dim=0)
