"""PopSpikeActor — host-side mirror of the reference's population-coded spiking actor.

Reference (four near-identical copies), cited from the AMP copy:
  AMP/AMP_for_hardware-main/rsl_rl/rsl_rl/modules/actor_critic.py
    :63-122   PopSpikeEncoderRegularSpike   (learnable Gaussian receptive fields, regular spikes)
    :190-271  SpikeMLP                      (current/voltage LIF layers, T timesteps)
    :125-150  PopSpikeDecoder               (grouped conv over each action's output population)
    :274-320  PopSpikeActor                 (obs -> (mu, exp(log_std)))

Same constructor arguments, same parameter names/shapes/initialisation (state-dict compatible), but the
modules below only HOLD parameters: forward and backward are the fused sm_100a kernels behind the C-ABI
(include/snn_b200.h), reached through one torch.autograd.Function.  There is no PyTorch fallback.
"""
from __future__ import annotations

import math
from typing import List, Sequence

import numpy as np
import torch
import torch.nn as nn

from .engine import SnnEngine


class PopSpikeEncoderRegularSpike(nn.Module):
    """Parameter holder for the encoder: mean, std of shape [1, obs_dim, pop_dim] (reference :85-92)."""

    def __init__(self, obs_dim: int, pop_dim: int, spike_ts: int, mean_range, std: float, device=None):
        super().__init__()
        self.obs_dim, self.pop_dim, self.spike_ts = obs_dim, pop_dim, spike_ts
        self.encoder_neuron_num = obs_dim * pop_dim
        self.device = device
        lo, hi = float(mean_range[0]), float(mean_range[1])
        delta = (hi - lo) / (pop_dim - 1)
        centres = torch.tensor([lo + delta * i for i in range(pop_dim)], dtype=torch.float32)
        self.mean = nn.Parameter(centres.reshape(1, 1, pop_dim).repeat(1, obs_dim, 1).contiguous())
        self.std = nn.Parameter(torch.zeros(1, obs_dim, pop_dim) + std)


class SpikeMLP(nn.Module):
    """Parameter holder for the LIF stack: hidden_layers[i], out_pop_layer are nn.Linear (reference :208-214)."""

    def __init__(self, in_pop_dim: int, out_pop_dim: int, hidden_sizes: Sequence[int], spike_ts: int, device=None):
        super().__init__()
        self.in_pop_dim, self.out_pop_dim = in_pop_dim, out_pop_dim
        self.hidden_sizes = list(hidden_sizes)
        self.hidden_num = len(hidden_sizes)
        self.spike_ts = spike_ts
        self.device = device
        dims = [in_pop_dim] + list(hidden_sizes)
        self.hidden_layers = nn.ModuleList([nn.Linear(dims[i], dims[i + 1]) for i in range(len(hidden_sizes))])
        self.out_pop_layer = nn.Linear(dims[-1], out_pop_dim)

    def linear_layers(self) -> List[nn.Linear]:
        return [*self.hidden_layers, self.out_pop_layer]


class PopSpikeDecoder(nn.Module):
    """Parameter holder for the decoder: Conv1d(act, act, pop, groups=act) weights [A,1,P], bias [A] (reference :137)."""

    def __init__(self, act_dim: int, pop_dim: int, output_activation=nn.Identity):
        super().__init__()
        self.act_dim, self.pop_dim = act_dim, pop_dim
        self.decoder = nn.Conv1d(act_dim, act_dim, pop_dim, groups=act_dim)
        self.output_activation = output_activation()


class _SpikingActorFn(torch.autograd.Function):
    """obs, params -> mu.  forward = snn_fwd_train / snn_fwd_infer, backward = snn_bwd."""

    @staticmethod
    def forward(ctx, engine: SnnEngine, L: int, grad_mode: bool, obs, enc_mean, enc_std, dec_w, dec_b, *wb):
        weights, biases = list(wb[:L]), list(wb[L:])
        # needs_input_grad mirrors requires_grad of the inputs whatever the caller's grad mode (and grad mode is always
        # off inside forward): grad_mode is torch.is_grad_enabled() at the call site.  Under no_grad / inference_mode
        # (act_inference, the rollout step) no trace is needed and the one-launch inference kernel runs.
        train = grad_mode and any(ctx.needs_input_grad)
        obs_c = obs.contiguous()
        mu, trace = engine.forward(obs_c, enc_mean, enc_std, weights, biases, dec_w, dec_b, train=train)
        if train:
            ctx.engine, ctx.L = engine, L
            ctx.save_for_backward(obs_c, mu, trace, enc_mean, enc_std, dec_w, dec_b, *wb)
        return mu

    @staticmethod
    def backward(ctx, g_mu):
        obs, mu, trace, enc_mean, enc_std, dec_w, dec_b, *wb = ctx.saved_tensors
        L = ctx.L
        weights, biases = list(wb[:L]), list(wb[L:])
        need_dobs = ctx.needs_input_grad[3]
        g = ctx.engine.backward(obs, mu, g_mu, trace, enc_mean, enc_std, weights, biases, dec_w, dec_b, need_dobs)
        return (None, None, None, g["obs"], g["enc_mean"], g["enc_std"], g["dec_w"], g["dec_b"], *g["weights"], *g["biases"])


class PopSpikeActor(nn.Module):
    """Population-coded spiking actor (reference :274-320).  forward(obs) -> (mu [B,A], std [A]).

    Extra keyword arguments select the per-fork deltas:
      enc_eps    1e-5 for the HUMANOID copy (pbrs-humanoid .../actor_critic.py:107)
      dec_tanh   True for the CASSIE copy (rsl_rl-master .../actor_critic.py:147)
      fwd_planes 3 = fp32-parity mode (weights re-encoded losslessly as three bf16 planes)
    """

    def __init__(self, obs_dim, act_dim, en_pop_dim, de_pop_dim, hidden_sizes, mean_range, std, spike_ts,
                 device=None, enc_eps: float = 0.0, dec_tanh: bool = False, fwd_planes: int = 3):
        super().__init__()
        self.encoder = PopSpikeEncoderRegularSpike(obs_dim, en_pop_dim, spike_ts, mean_range, std, device)
        self.snn = SpikeMLP(obs_dim * en_pop_dim, act_dim * de_pop_dim, hidden_sizes, spike_ts, device)
        self.decoder = PopSpikeDecoder(act_dim, de_pop_dim)
        self.log_std = nn.Parameter(torch.as_tensor(-0.001 * np.ones(act_dim, dtype=np.float32)))
        self.obs_dim, self.act_dim = obs_dim, act_dim
        self.spike_ts = spike_ts
        self.enc_eps, self.dec_tanh, self.fwd_planes = enc_eps, dec_tanh, fwd_planes
        self._engine = SnnEngine(obs_dim, act_dim, en_pop_dim, de_pop_dim, list(hidden_sizes), spike_ts,
                                 dec_tanh=dec_tanh, enc_eps=enc_eps, fwd_planes=fwd_planes)

    @property
    def engine(self) -> SnnEngine:
        return self._engine

    def __deepcopy__(self, memo):   # export path deep-copies .actor (legged_gym/utils/helpers.py:188)
        new = PopSpikeActor(self.obs_dim, self.act_dim, self.encoder.pop_dim, self.decoder.pop_dim,
                            self.snn.hidden_sizes, (-5, 5), math.sqrt(0.15), self.spike_ts,
                            enc_eps=self.enc_eps, dec_tanh=self.dec_tanh, fwd_planes=self.fwd_planes)
        new.to(self.log_std.device)
        new.load_state_dict(self.state_dict())
        return new

    def mean_only(self, obs: torch.Tensor) -> torch.Tensor:
        layers = self.snn.linear_layers()
        return _SpikingActorFn.apply(self._engine, len(layers), torch.is_grad_enabled(), obs, self.encoder.mean, self.encoder.std,
                                     self.decoder.decoder.weight, self.decoder.decoder.bias,
                                     *[l.weight for l in layers], *[l.bias for l in layers])

    def forward(self, obs: torch.Tensor):
        mu = self.mean_only(obs)
        std = torch.exp(self.log_std)
        return mu, std
