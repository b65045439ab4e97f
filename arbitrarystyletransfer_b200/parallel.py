"""Multi-GPU plumbing for the hot path: one process per GPU, torch.distributed for rendezvous.

The reference has no distributed code at all.  The path shards naturally:
every op is per-sample, so
  * inference (BASELINE config 4) splits the batch contiguously across ranks, replicates the
    weights and needs NO collective;
  * training all-reduces ONE flat fp32 bucket holding every trainable gradient (14.0 MB for the
    classic decoder) per step -- NCCL over NVLink/NVSwitch on GPUs, gloo in the CPU tests -- and
    scales by 1/world so that per-shard 'mean' losses average exactly like the single-process
    reference step on the global batch (train.py:287-300 then runs unchanged on identical replicas).
"""
from __future__ import annotations

from typing import Iterable, Sequence

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [lo, hi) of `n_items` owned by `rank`; remainders go to the first ranks."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(x: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    lo, hi = shard_range(x.shape[0], rank, world)
    return x[lo:hi]


class GradBucket:
    """All trainable gradients as views into one flat fp32 buffer, in reverse parameter order
    (the order backward produces them), so a step costs a single collective."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel, device=dev, dtype=torch.float32)
        off = 0
        for p in reversed(self.params):
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)   # .grad aliases the bucket
            off += n

    def zero(self):
        self.flat.zero_()

    def all_reduce_mean(self, group=None):
        """Sum over ranks, divide by world size.  No-op without an initialised process group."""
        if not (dist.is_available() and dist.is_initialized()):
            return self.flat
        world = dist.get_world_size(group)
        if world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            self.flat.div_(world)
        return self.flat


def broadcast_parameters(params: Sequence[torch.Tensor], src: int = 0, group=None) -> None:
    """Make replicas identical before the first step (weights are replicated, never sharded)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    for p in params:
        dist.broadcast(p.data if isinstance(p, torch.nn.Parameter) else p, src=src, group=group)


def gather_shards(local: torch.Tensor, n_items: int, group=None) -> torch.Tensor | None:
    """Collect per-rank output shards on rank 0 (used by tests / demos only: the inference data
    path itself has no collective)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_range(n_items, r, world) for r in range(world)]
    maxn = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((maxn,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
    dist.gather(pad, bufs, dst=0, group=group)
    if rank != 0:
        return None
    return torch.cat([b[: hi - lo] for b, (lo, hi) in zip(bufs, sizes)], 0)
