"""Multi-GPU plumbing for the batch-sharded forward: one process per GPU, samples are independent,
weights / graphs are replicated, and the only communication is ONE all-gather of the per-rank outputs at the loss
boundary.  torch.distributed (NCCL over NVLink on the GPUs, gloo in the CPU tests) is the transport."""

from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(total_batch: int, rank: int, world: int):
    """Contiguous, balanced [start, stop) of the global batch owned by `rank` (earlier ranks take the remainder)."""
    base, rem = divmod(total_batch, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def all_gather_batch(y: torch.Tensor, total_batch: int, group=None) -> torch.Tensor:
    """Gathers per-rank outputs [b_r, ...] into [total_batch, ...] in rank order; shards may differ by one sample."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_range(total_batch, r, world) for r in range(world)]
    assert y.shape[0] == sizes[rank][1] - sizes[rank][0], "local shard does not match shard_range"
    bmax = max(b - a for a, b in sizes)
    if all(b - a == bmax for a, b in sizes):
        out = torch.empty((world * bmax,) + tuple(y.shape[1:]), dtype=y.dtype, device=y.device)
        dist.all_gather_into_tensor(out, y.contiguous(), group=group)
        return out
    pad = torch.zeros((bmax,) + tuple(y.shape[1:]), dtype=y.dtype, device=y.device)
    pad[: y.shape[0]] = y
    buf = torch.empty((world * bmax,) + tuple(y.shape[1:]), dtype=y.dtype, device=y.device)
    dist.all_gather_into_tensor(buf, pad, group=group)
    return torch.cat([buf[r * bmax : r * bmax + (b - a)] for r, (a, b) in enumerate(sizes)], dim=0)


def max_over_ranks(value: float, device, group=None) -> float:
    t = torch.tensor([float(value)], device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
