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


class BucketedGradExchange:
    """DDP's gradient step without DDP: a flat fp32 gradient buffer of `n_params` elements cut into `bucket_bytes` buckets
    (torch DDP's default is 25 MB), each all-reduced asynchronously on NCCL's stream.  ``step(grads)`` copies the given
    gradient tensors into the head of the buffer (the rest stands for the other parameters of the model the scan sits in),
    waits for the PREVIOUS step's buckets (they overlapped this step's compute, as DDP overlaps buckets with the rest of
    the backward) and launches this step's; ``wait=True`` also completes them (last step / optimizer boundary).

    ``buckets_per_step=None``: every bucket every step (one whole model's gradients per call).  ``buckets_per_step=k``: the
    next k buckets in round-robin order, i.e. the gradient set leaves bucket by bucket over ``ceil(n_buckets / k)`` calls --
    DDP's schedule when one call is ONE layer's backward of a model with that many layers (a bucket becomes ready when the
    layers that fill it are done and travels while the next layers' backward runs)."""

    def __init__(self, n_params: int, device, bucket_bytes: int = 25 << 20, group=None, buckets_per_step: Optional[int] = None):
        self.flat = torch.zeros(n_params, dtype=torch.float32, device=device)
        per = max(1, bucket_bytes // 4)
        self.buckets = [self.flat[i:i + per] for i in range(0, n_params, per)]
        self.group, self.pending = group, []
        self.model_bytes = n_params * 4
        self.buckets_per_step = None if buckets_per_step is None else max(1, min(int(buckets_per_step), len(self.buckets)))
        self.cursor = 0
        self.steps = 0
        self.bytes_sent = 0

    @property
    def bytes_per_step(self) -> float:
        """Gradient bytes all-reduced per step() call (average over a full round of the buckets)."""
        if self.buckets_per_step is None:
            return float(self.model_bytes)
        return self.model_bytes * self.buckets_per_step / len(self.buckets)

    def drain(self):
        for h in self.pending:
            h.wait()
        self.pending.clear()

    def _next_buckets(self):
        if self.buckets_per_step is None:
            return list(self.buckets)
        n = len(self.buckets)
        sel = [self.buckets[(self.cursor + i) % n] for i in range(self.buckets_per_step)]
        self.cursor = (self.cursor + self.buckets_per_step) % n
        return sel

    def step(self, grads, wait: bool = False):
        self.drain()
        off = 0
        for g in grads:
            if g is None:
                continue
            n = g.numel()
            self.flat[off:off + n].copy_(g.reshape(-1))
            off += n
        sel = self._next_buckets()
        if dist.is_initialized() and dist.get_world_size(self.group) > 1:
            self.pending = [dist.all_reduce(b, group=self.group, async_op=True) for b in sel]
        self.bytes_sent += sum(b.numel() * 4 for b in sel)
        self.steps += 1
        if wait:
            self.drain()

    def measure_alone(self, device, reps: int = 3):
        """Time the all-reduce of the WHOLE gradient set, bucket by bucket, with nothing else on the GPU (max over ranks):
        bus GB/s = 2 (n-1)/n x bytes / time."""
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        if world < 2:
            return {}
        self.drain()
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            for h in [dist.all_reduce(b, group=self.group, async_op=True) for b in self.buckets]:
                h.wait()
        e1.record()
        torch.cuda.synchronize()
        ms = max_over_ranks(e0.elapsed_time(e1) / reps, device)
        return {"alone_ms": ms, "alone_bus_gbs": 2.0 * (world - 1) / world * self.model_bytes / (ms * 1e-3) / 1e9}

    def report(self):
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        n = len(self.buckets)
        every = 1 if self.buckets_per_step is None else -(-n // self.buckets_per_step)
        return {"model_bytes": self.model_bytes, "buckets": n, "bucket_bytes": self.buckets[0].numel() * 4,
                "buckets_per_step": n if self.buckets_per_step is None else self.buckets_per_step,
                "steps_per_gradient_set": every, "bytes_per_step": self.bytes_per_step, "world": world,
                "note": "all-reduce bus traffic per rank = 2 (n-1)/n x bytes_per_step; overlapped with the next step's kernels"}
