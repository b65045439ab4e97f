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
    (the order backward produces them), so a step costs a single collective.

    The aliasing ``p.grad is a view of self.flat`` is what makes the single collective correct, and the
    reference's own step breaks it: ``optim.zero_grad()`` (train.py:287, train_autoencoder.py:140) defaults to
    ``set_to_none=True`` in torch >= 2.0, after which backward allocates fresh ``.grad`` tensors.  ``all_reduce_mean``
    therefore re-checks every parameter and *adopts* foreign gradients (copies them into the bucket and re-installs
    the view), so the reference step runs unchanged; ``bucket.zero()`` / ``zero_grad(set_to_none=False)`` avoid
    that copy."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel, device=dev, dtype=torch.float32)
        self._views = []
        off = 0
        for p in reversed(self.params):
            n = p.numel()
            v = self.flat[off:off + n].view_as(p)
            self._views.append((p, v))
            p.grad = v                                   # .grad aliases the bucket
            off += n

    def zero(self):
        self.flat.zero_()
        self.adopt()

    def aliased(self) -> bool:
        """True when every parameter's ``.grad`` still is its bucket view."""
        return all(p.grad is not None and p.grad.data_ptr() == v.data_ptr() and p.grad.shape == v.shape
                   for p, v in self._views)

    def adopt(self) -> int:
        """Re-install the bucket views.  A parameter whose ``.grad`` was replaced (``zero_grad(set_to_none=True)``
        followed by backward) has its gradient copied into the bucket first; a parameter without a gradient
        contributes zeros.  Returns how many views had to be repaired."""
        fixed = 0
        for p, v in self._views:
            g = p.grad
            if g is not None and g.data_ptr() == v.data_ptr() and g.shape == v.shape:
                continue
            if g is None:
                v.zero_()
            else:
                v.copy_(g)
            p.grad = v
            fixed += 1
        return fixed

    def all_reduce_mean(self, group=None, local_count: int | None = None, async_op: bool = False):
        """Sum over ranks and divide: by the world size (equal shards, per-shard 'mean' losses), or -- when
        ``local_count`` (this rank's number of samples) is given -- weighted by shard size, so that unequal
        shards (``shard_range`` with N % world != 0) still give the global-batch mean.  No-op without an initialised
        process group.  With ``async_op`` the collective's work handle is returned (the division is folded into
        the pre-scale, so nothing runs after the wait)."""
        self.adopt()
        if not (dist.is_available() and dist.is_initialized()):
            return self.flat
        world = dist.get_world_size(group)
        if world == 1:
            return self.flat
        if local_count is None:
            self.flat.div_(world)
        else:
            tot = torch.tensor([float(local_count)], device=self.flat.device)
            dist.all_reduce(tot, op=dist.ReduceOp.SUM, group=group)
            self.flat.mul_(float(local_count) / tot)
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        return work if async_op else self.flat


def broadcast_parameters(params: Sequence[torch.Tensor], src: int = 0, group=None) -> None:
    """Make replicas identical before the first step (weights are replicated, never sharded)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    for p in params:
        dist.broadcast(p.data if isinstance(p, torch.nn.Parameter) else p, src=src, group=group)


def broadcast_module(module: torch.nn.Module, src: int = 0, group=None) -> None:
    """Parameters AND buffers (BatchNorm running statistics, ``num_batches_tracked``) from ``src``: replicas start
    identical, and a checkpoint written by any rank after ``broadcast_module`` equals rank ``src``'s
    (per-shard BatchNorm statistics diverge per rank; broadcast before saving)."""
    broadcast_parameters(list(module.parameters()) + list(module.buffers()), src, group)


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
