"""Data-parallel plumbing around the scan (the only multi-GPU pattern the reference uses: DDP).

The scan is per image and never shards; ranks process disjoint batches and exchange ONLY parameter gradients
(dA, dD, d_delta_bias here; main_pretrain.py:167-169 wraps the model in DistributedDataParallel for the same effect).
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


def allreduce_param_grads(grads: Iterable[Optional[torch.Tensor]], group=None, average: bool = True) -> List[Optional[torch.Tensor]]:
    """One bucketed all-reduce (flatten -> all_reduce -> unflatten) of the given gradient tensors, in place.
    average=True divides by the world size like DDP does."""
    tensors = [g for g in grads if g is not None]
    if not tensors or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return list(grads)
    flat = torch.cat([t.reshape(-1).float() for t in tensors])
    dist.all_reduce(flat, group=group)
    if average:
        flat /= dist.get_world_size(group)
    off = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[off:off + n].view_as(t).to(t.dtype))
        off += n
    return list(grads)


def max_over_ranks(value: float, device=None, group=None) -> float:
    """Timing rule of the bench contract: a step takes as long as its slowest rank."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def shard_batch(global_batch: int, rank: int, world: int) -> range:
    """Images [lo, hi) of a global batch owned by `rank` (contiguous, remainder to the first ranks)."""
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return range(lo, lo + base + (1 if rank < rem else 0))
