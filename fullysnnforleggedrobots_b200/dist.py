"""Data parallelism for the PPO update: one process per GPU, environments and minibatches sharded by rank,
parameters replicated, and ONE all-reduce (sum) of the flattened actor-critic gradient per minibatch step
(SURVEY.md §8e).  The reference has no distributed code; with world_size == 1 everything here is a no-op and
the update is the single-process algorithm.

NCCL over NVLink on the GPU box; the same code runs under gloo on CPU for the world_size-2 tests.
"""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None) -> "GradientAllReducer | None":
    """Initialise torch.distributed from RANK / WORLD_SIZE / MASTER_* (torchrun) and return a reducer,
    or None when running single-process."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return None
    if not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group(backend=backend)
    return GradientAllReducer()


def shard_range(n: int, rank: int, world: int):
    """Contiguous block of environments owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def normalize_advantages_global(adv: torch.Tensor, allreduce_sum) -> torch.Tensor:
    """(adv - mean) / (std + 1e-8) with mean / unbiased std over ALL ranks' samples
    (reference rollout_storage.py:137-138 computes them over the single process's whole rollout).
    `allreduce_sum(t)` must sum a float64 tensor over ranks in place and return it."""
    a = adv.double()
    stats = torch.stack([a.sum(), (a * a).sum(), torch.tensor(float(a.numel()), device=a.device, dtype=torch.float64)])
    stats = allreduce_sum(stats)
    n = stats[2]
    mean = stats[0] / n
    var = (stats[1] - n * mean * mean) / (n - 1)
    return (adv - mean.to(adv.dtype)) / (var.clamp_min(0).sqrt().to(adv.dtype) + 1e-8)


class GradientAllReducer:
    def __init__(self, group=None):
        self.group = group
        self.world_size = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self._flat = None

    def broadcast_parameters(self, module: torch.nn.Module, src: int = 0):
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src=src, group=self.group)

    def allreduce_sum(self, t: torch.Tensor) -> torch.Tensor:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def allreduce_gradients(self, module: torch.nn.Module):
        """grad <- mean over ranks, through one flat bucket (one collective per minibatch step)."""
        params = [p for p in module.parameters() if p.requires_grad]
        n = sum(p.numel() for p in params)
        dev = params[0].device
        if self._flat is None or self._flat.numel() != n or self._flat.device != dev:
            self._flat = torch.zeros(n, dtype=torch.float32, device=dev)
        flat = self._flat
        off = 0
        for p in params:
            k = p.numel()
            if p.grad is None:
                flat[off:off + k].zero_()
            else:
                flat[off:off + k].copy_(p.grad.reshape(-1))
            off += k
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        flat.div_(self.world_size)
        off = 0
        for p in params:
            k = p.numel()
            if p.grad is not None:
                p.grad.copy_(flat[off:off + k].view_as(p.grad))
            off += k
