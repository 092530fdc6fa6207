"""Data parallelism for the PPO update: one process per GPU, environments and minibatches sharded by rank,
parameters replicated, and ONE all-reduce (sum) of the flattened actor-critic gradient per minibatch step
(SURVEY.md §8e).  The reference has no distributed code; with world_size == 1 everything here is a no-op and
the update is the single-process algorithm.

NCCL over NVLink on the GPU box; the same code runs under gloo on CPU for the world_size-2 tests.

r2: on CUDA the per-step all-reduce is not an NCCL call any more but part of the optimizer kernel (`P2PGradientSum`,
snn_clip_adam_kl_p2p: every rank reads all ranks' gradient buckets over NVLink peer memory and sums them in rank order);
NCCL / gloo keep the rendezvous, the parameter broadcast and the once-per-iteration advantage statistics.
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


class P2PGradientSum:
    """Peer-mapped staging buffers + flag arrays of all ranks for snn_clip_adam_kl_p2p (one node, 2..8 ranks).
    Every rank cudaMallocs its own buffers inside the library, the CUDA IPC handles travel through
    `all_gather_object`, every rank opens the others'.  Raises when any rank cannot set it up (the caller falls back to
    the NCCL all-reduce on every rank alike)."""

    def __init__(self, n_floats: int, group=None):
        import ctypes as C
        from . import _lib
        L = _lib.lib()
        self.world_size, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if not 2 <= self.world_size <= 8:
            raise RuntimeError("peer-memory gradient sum: 2..8 ranks")
        self.stride = (int(n_floats) + 255) // 256 * 256
        own, handles, err = [], [], None
        try:
            for nbytes in (2 * self.stride * 4, 256):
                ptr, h = C.c_void_p(), C.create_string_buffer(64)
                _lib.check(L.snn_p2p_alloc(nbytes, C.byref(ptr), h), "snn_p2p_alloc")
                own.append(ptr.value)
                handles.append(h.raw)
        except Exception as e:  # noqa: BLE001   (reported to every rank below)
            err = str(e)
        everyone = [None] * self.world_size
        dist.all_gather_object(everyone, (handles, err), group=group)
        bad = [e for _, e in everyone if e is not None]
        self._own, self._opened = own, []
        if bad:
            self.close()
            raise RuntimeError("peer-memory gradient sum unavailable: " + bad[0])
        stage, flags = (C.c_void_p * self.world_size)(), (C.c_void_p * self.world_size)()
        ok = True
        try:
            for q, (hs, _) in enumerate(everyone):
                if q == self.rank:
                    stage[q], flags[q] = own[0], own[1]
                    continue
                for k, arr in ((0, stage), (1, flags)):
                    ptr = C.c_void_p()
                    _lib.check(L.snn_p2p_open(hs[k], C.byref(ptr)), "snn_p2p_open")
                    self._opened.append(ptr.value)
                    arr[q] = ptr.value
        except Exception as e:  # noqa: BLE001
            ok, err = False, str(e)
        flags_ok = [None] * self.world_size
        dist.all_gather_object(flags_ok, ok, group=group)
        if not all(flags_ok):
            self.close()
            raise RuntimeError("peer-memory gradient sum unavailable: " + (err or "a peer could not map the buffers"))
        self.stage, self.flags = stage, flags
        self.epoch = torch.zeros(1, dtype=torch.int32, device="cuda")

    def close(self):
        from . import _lib
        L = _lib.lib()
        for p in getattr(self, "_opened", []):
            L.snn_p2p_close(p)
        for p in getattr(self, "_own", []):
            L.snn_p2p_free(p)
        self._opened, self._own = [], []


class GradientAllReducer:
    def __init__(self, group=None):
        self.group = group
        self.world_size = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self._flat = None
        self.p2p = None             # P2PGradientSum once enable_p2p() succeeded

    def enable_p2p(self, n_floats: int) -> bool:
        """Set the peer-memory gradient sum up for a bucket of n_floats (all ranks call this together).  False (and a
        note in `p2p_note`) when it is unavailable: not CUDA, more than 8 ranks, SNN_DP_P2P=0, IPC mapping failed."""
        self.p2p_note = ""
        if self.p2p is not None and self.p2p.stride >= n_floats:
            return True
        if os.environ.get("SNN_DP_P2P", "1") == "0" or not torch.cuda.is_available() or dist.get_backend(self.group) != "nccl":
            self.p2p_note = "disabled (SNN_DP_P2P=0 or not a CUDA / NCCL job)"
            return False
        try:
            self.p2p = P2PGradientSum(n_floats, self.group)
            return True
        except RuntimeError as e:
            self.p2p, self.p2p_note = None, str(e)
            return False

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
