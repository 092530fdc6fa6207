"""FlatAdam — Adam + clip_grad_norm_ over ONE flat fp32 buffer, driven by snn_clip_adam (one kernel pair per step).

Replaces `nn.utils.clip_grad_norm_` + `torch.optim.Adam.step()` of the reference update
(AMP/AMP_for_hardware-main/rsl_rl/rsl_rl/algorithms/ppo.py:176-177).  Parameters of the module become views into
`flat_params`, their `.grad` views into `flat_grads`, so that
  * the actor's BPTT kernels write gradients straight into the bucket that NCCL all-reduces (no pack/unpack),
  * grad-norm, clip and the Adam update are two launches regardless of the number of tensors,
  * the learning rate and step counter live on the device (no host round trip; CUDA-graph capturable).
It is still a torch.optim.Optimizer: state_dict()/load_state_dict() use torch.optim.Adam's layout
(per-parameter 'step', 'exp_avg', 'exp_avg_sq'), which is what the reference runner checkpoints.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, List

import torch

from . import _lib


class FlatAdam(torch.optim.Optimizer):
    def __init__(self, params: Iterable[torch.nn.Parameter], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0,
                 tail: int = 0):
        params = [p for p in params]
        if not params:
            raise ValueError("no parameters")
        dev = params[0].device
        if dev.type != "cuda":
            raise RuntimeError("FlatAdam drives CUDA kernels: parameters must live on a CUDA device")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.n = sum(p.numel() for p in params)
        self.tail = tail                      # extra floats after the gradients (KL statistic for the all-reduce)
        self.flat_params = torch.empty(self.n, dtype=torch.float32, device=dev)
        self.flat_grads = torch.zeros(self.n + tail, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(self.n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(self.n, dtype=torch.float32, device=dev)
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.lr_dev = torch.full((1,), float(lr), dtype=torch.float32, device=dev)
        self.norm_dev = torch.zeros(1, dtype=torch.float32, device=dev)
        self._scratch = torch.empty(4096, dtype=torch.uint8, device=dev)
        self.offsets: List[int] = []
        off = 0
        with torch.no_grad():
            for p in params:
                k = p.numel()
                self.flat_params[off:off + k].copy_(p.data.reshape(-1))
                p.data = self.flat_params[off:off + k].view(p.shape)
                p.grad = self.flat_grads[off:off + k].view(p.shape)
                self.state[p] = {"step": self.step_dev, "exp_avg": self.exp_avg[off:off + k].view(p.shape),
                                 "exp_avg_sq": self.exp_avg_sq[off:off + k].view(p.shape)}
                self.offsets.append(off)
                off += k
        self.params = params

    def zero_grad(self, set_to_none: bool = False):   # grads are views of the flat bucket: never drop them
        self.flat_grads.zero_()

    def set_lr(self, lr: float):
        self.lr_dev.fill_(float(lr))
        for g in self.param_groups:
            g["lr"] = float(lr)

    def sync_lr_from_device(self) -> float:
        lr = float(self.lr_dev.item())
        for g in self.param_groups:
            g["lr"] = lr
        return lr

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        # copy loaded moments back into the flat buffers and re-alias
        with torch.no_grad():
            for p, off in zip(self.params, self.offsets):
                st = self.state.get(p, {})
                k = p.numel()
                if "exp_avg" in st:
                    self.exp_avg[off:off + k].copy_(st["exp_avg"].reshape(-1))
                    self.exp_avg_sq[off:off + k].copy_(st["exp_avg_sq"].reshape(-1))
                    step = st.get("step", 0)
                    self.step_dev.fill_(int(step.item()) if torch.is_tensor(step) else int(step))
                self.state[p] = {"step": self.step_dev, "exp_avg": self.exp_avg[off:off + k].view(p.shape),
                                 "exp_avg_sq": self.exp_avg_sq[off:off + k].view(p.shape)}
        self.set_lr(self.param_groups[0]["lr"])

    @torch.no_grad()
    def step_p2p(self, p2p, max_grad_norm: float = 0.0, grad_scale: float = 1.0, adaptive: bool = False,
                 kl_scale: float = 1.0, desired_kl: float = 0.0):
        """Data-parallel step: the sum of the whole bucket (gradients + the KL statistic in its tail) over the ranks, the
        adaptive-KL rule on the summed statistic, clip and Adam in ONE kernel over peer memory (dist.P2PGradientSum)."""
        g = self.param_groups[0]
        b1, b2 = g["betas"]
        dev = self.flat_params.device
        ntot = self.flat_grads.numel()
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().snn_clip_adam_kl_p2p(
                C.c_void_p(self.flat_params.data_ptr()), C.c_void_p(self.flat_grads.data_ptr()),
                C.c_void_p(self.exp_avg.data_ptr()), C.c_void_p(self.exp_avg_sq.data_ptr()), self.n, ntot,
                C.c_void_p(self.lr_dev.data_ptr()), C.c_void_p(self.step_dev.data_ptr()), float(b1), float(b2),
                float(g["eps"]), float(g["weight_decay"]), float(max_grad_norm), float(grad_scale),
                self.n if adaptive else -1, float(kl_scale), float(desired_kl),
                C.c_void_p(self.norm_dev.data_ptr()), C.c_void_p(self._scratch.data_ptr()), p2p.stage, p2p.flags,
                p2p.rank, p2p.world_size, p2p.stride, C.c_void_p(p2p.epoch.data_ptr()),
                C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "snn_clip_adam_kl_p2p")

    @torch.no_grad()
    def step(self, max_grad_norm: float = 0.0, grad_scale: float = 1.0, closure=None, kl=None, kl_scale: float = 1.0,
             desired_kl: float = 0.0):
        """clip (if max_grad_norm > 0) + Adam on the flat buffers; lr is read from lr_dev on the device.
        kl (a 1-element device tensor): the adaptive-KL schedule of PPO (ppo.py:145-148) is applied to lr_dev first,
        inside the same kernel (kl_mean = kl[0] * kl_scale)."""
        g = self.param_groups[0]
        b1, b2 = g["betas"]
        dev = self.flat_params.device
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().snn_clip_adam_kl(
                C.c_void_p(self.flat_params.data_ptr()), C.c_void_p(self.flat_grads.data_ptr()),
                C.c_void_p(self.exp_avg.data_ptr()), C.c_void_p(self.exp_avg_sq.data_ptr()), self.n,
                C.c_void_p(self.lr_dev.data_ptr()), C.c_void_p(self.step_dev.data_ptr()), float(b1), float(b2),
                float(g["eps"]), float(g["weight_decay"]), float(max_grad_norm), float(grad_scale),
                C.c_void_p(kl.data_ptr()) if kl is not None else None, float(kl_scale), float(desired_kl),
                C.c_void_p(self.norm_dev.data_ptr()), C.c_void_p(self._scratch.data_ptr()),
                C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "snn_clip_adam_kl")
