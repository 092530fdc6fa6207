"""CPU restatement of one PPO minibatch step of the reference, on top of the spiking-actor oracle.

TEST INFRASTRUCTURE ONLY (see snn_oracle.py's header for who may import this).

Follows AMP/AMP_for_hardware-main/rsl_rl/rsl_rl/algorithms/ppo.py:120-187 ("PPO:<line>") and
.../modules/actor_critic.py:398-430 ("AC:<line>"): act() -> Normal(mu, mu*0+std), log-prob / entropy sums,
dense ELU critic, adaptive-KL learning rate, clipped surrogate + clipped value loss, backward,
clip_grad_norm_, Adam.  Parity pinning: tests/golden/ppo_step_*.npz, produced by running the reference
PPO + ActorCritic themselves (tests/golden/make_golden.py).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List

import torch

from . import snn_oracle as O


@dataclass
class PPOConfig:           # AMP legged_robot_config.py:226-238 defaults
    clip_param: float = 0.2
    value_loss_coef: float = 1.0
    entropy_coef: float = 0.01
    learning_rate: float = 1e-3
    max_grad_norm: float = 1.0
    use_clipped_value_loss: bool = True
    schedule: str = "adaptive"
    desired_kl: float = 0.01


class OracleActorCritic:
    """Leaf-tensor actor-critic built from a reference-style state dict (keys 'actor.*', 'critic.*', 'std')."""

    def __init__(self, sd: Dict[str, torch.Tensor], spike_ts=5, enc_eps=0.0, dec_tanh=False):
        self.names = list(sd.keys())
        self.t = {k: v.detach().clone().float().requires_grad_(True) for k, v in sd.items()}
        self.spike_ts, self.enc_eps, self.dec_tanh = spike_ts, enc_eps, dec_tanh

    def parameters(self) -> List[torch.Tensor]:
        return [self.t[k] for k in self.names]

    def actor_params(self) -> O.ActorParams:
        t = self.t
        ws, bs, i = [], [], 0
        while f"actor.snn.hidden_layers.{i}.weight" in t:
            ws.append(t[f"actor.snn.hidden_layers.{i}.weight"]); bs.append(t[f"actor.snn.hidden_layers.{i}.bias"]); i += 1
        ws.append(t["actor.snn.out_pop_layer.weight"]); bs.append(t["actor.snn.out_pop_layer.bias"])
        return O.ActorParams(t["actor.encoder.mean"], t["actor.encoder.std"], ws, bs, t["actor.decoder.decoder.weight"],
                             t["actor.decoder.decoder.bias"], t["actor.log_std"], spike_ts=self.spike_ts,
                             enc_eps=self.enc_eps, dec_tanh=self.dec_tanh)

    def critic(self, x):
        ws, bs, i = [], [], 0
        while f"critic.{i}.weight" in self.t:
            ws.append(self.t[f"critic.{i}.weight"]); bs.append(self.t[f"critic.{i}.bias"]); i += 2
        return O.critic_forward(x, ws, bs)


def ppo_minibatch_losses(ac: OracleActorCritic, batch, cfg: PPOConfig):
    """PPO:131-171.  batch = (obs, critic_obs, actions, target_values, advantages, returns, old_logp, old_mu, old_sigma)."""
    obs, critic_obs, actions, target_values, advantages, returns, old_logp, old_mu, old_sigma = batch
    mu, std = O.actor_forward_autograd(obs, ac.actor_params())
    sigma = mu * 0. + std                                   # AC:414
    logp = O.normal_log_prob(actions, mu, sigma)            # AC:420-421
    value = ac.critic(critic_obs)                           # AC:428-430
    entropy = O.normal_entropy(mu, sigma)                   # AC:406-408
    with torch.no_grad():
        kl_mean = O.kl_gaussian(mu, sigma, old_mu, old_sigma).mean()    # PPO:141-143
    loss, sur, vl = O.ppo_losses(logp, old_logp, advantages, value, target_values, returns, entropy,
                                 clip=cfg.clip_param, value_coef=cfg.value_loss_coef, ent_coef=cfg.entropy_coef,
                                 clipped_value=cfg.use_clipped_value_loss)
    return loss, sur, vl, kl_mean


def adapt_lr(lr: float, kl_mean: float, cfg: PPOConfig) -> float:
    """PPO:145-148."""
    if cfg.schedule != "adaptive" or cfg.desired_kl is None:
        return lr
    if kl_mean > cfg.desired_kl * 2.0:
        return max(1e-5, lr / 1.5)
    if cfg.desired_kl / 2.0 > kl_mean > 0.0:
        return min(1e-2, lr * 1.5)
    return lr


def ppo_minibatch_step(ac: OracleActorCritic, optimizer: torch.optim.Optimizer, batch, cfg: PPOConfig, lr: float):
    """One iteration of the loop body PPO:131-180.  Returns dict of scalars and the new learning rate."""
    loss, sur, vl, kl_mean = ppo_minibatch_losses(ac, batch, cfg)
    lr = adapt_lr(lr, float(kl_mean), cfg)
    for g in optimizer.param_groups:
        g["lr"] = lr
    optimizer.zero_grad()
    loss.backward()
    gn = torch.nn.utils.clip_grad_norm_([p for p in ac.parameters()], cfg.max_grad_norm)
    optimizer.step()
    return {"loss": float(loss.detach()), "surrogate": float(sur.detach()), "value_loss": float(vl.detach()), "kl": float(kl_mean),
            "grad_norm": float(gn)}, lr
