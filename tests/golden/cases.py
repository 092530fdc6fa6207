"""Deterministic parity cases shared by make_golden.py (which runs the reference on them in the
build container) and by the tests (which run the oracle / the CUDA path on them anywhere).

Inputs and weights are regenerated from seeds with the CPU torch generator, so only the
reference's OUTPUTS need to be committed as fixtures.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List

import torch


@dataclass(frozen=True)
class Case:
    name: str
    variant: str            # "amp" | "cassie" | "humanoid" | "rma"
    obs_dim: int            # actor input dim (RMA: obs + latent, fed directly to the actor)
    hidden: tuple
    act_dim: int
    batch: int
    spike_ts: int = 5
    pop: int = 10
    obs_kind: str = "normal"   # "normal" | "stress"
    seed: int = 0
    trained: bool = False      # perturb weights away from default init (as after a few Adam steps)
    full: bool = True          # store full traces / grads (small cases) or summaries (big cases)

    @property
    def enc_eps(self) -> float:
        return 1e-5 if self.variant == "humanoid" else 0.0

    @property
    def dec_tanh(self) -> bool:
        return self.variant == "cassie"


CASES: List[Case] = [
    Case("tiny_amp", "amp", 6, (32, 24), 3, 16, seed=1),
    Case("tiny_cassie_tanh", "cassie", 5, (16,), 2, 12, seed=2),
    Case("tiny_humanoid_eps", "humanoid", 7, (24, 24, 24), 4, 10, seed=3),
    Case("tiny_rma_T8", "rma", 9, (40, 16), 3, 8, spike_ts=8, seed=4),
    Case("tiny_stress", "amp", 6, (32, 24), 3, 64, obs_kind="stress", seed=5, trained=True),
    Case("ragged_b1", "amp", 3, (8,), 1, 1, seed=6),
    Case("ragged_b131", "cassie", 11, (72, 40), 5, 131, seed=7, trained=True),
    # BASELINE.json configs[0] dims (C1), batch cut to keep the fixture small
    Case("c1_dims", "amp", 48, (256, 256), 12, 256, seed=8, full=False),
    Case("humanoid_dims", "humanoid", 38, (256, 256, 256), 10, 128, seed=9, full=False),
    Case("amp_actual_dims", "amp", 42, (512, 256, 128), 12, 128, seed=10, full=False, trained=True),
]


def case_by_name(name: str) -> Case:
    for c in CASES:
        if c.name == name:
            return c
    raise KeyError(name)


def _uniform(g: torch.Generator, shape, bound: float) -> torch.Tensor:
    return (torch.rand(shape, generator=g, dtype=torch.float32) * 2.0 - 1.0) * bound


def make_state_dict(c: Case) -> Dict[str, torch.Tensor]:
    """Actor state dict with the reference's key names and nn.Linear-style
    U(+-1/sqrt(fan_in)) init, drawn from our own generator (NOT torch's global RNG)."""
    g = torch.Generator().manual_seed(1000 + c.seed)
    O, P, A = c.obs_dim, c.pop, c.act_dim
    sd: Dict[str, torch.Tensor] = {}
    mean = torch.linspace(-5.0, 5.0, P, dtype=torch.float32).reshape(1, 1, P).repeat(1, O, 1)
    std = torch.full((1, O, P), math.sqrt(0.15), dtype=torch.float32)
    if c.trained:
        mean = mean + _uniform(g, mean.shape, 0.05)
        std = std + _uniform(g, std.shape, 0.02)
    sd["actor.encoder.mean"] = mean
    sd["actor.encoder.std"] = std
    dims = [O * P] + list(c.hidden) + [A * P]
    for i in range(len(dims) - 1):
        k, n = dims[i], dims[i + 1]
        bound = 1.0 / math.sqrt(k)
        scale = 1.6 if c.trained else 1.0
        name = f"actor.snn.hidden_layers.{i}" if i < len(c.hidden) else "actor.snn.out_pop_layer"
        sd[name + ".weight"] = _uniform(g, (n, k), bound * scale)
        sd[name + ".bias"] = _uniform(g, (n,), bound * scale)
    bound = 1.0 / math.sqrt(P)
    sd["actor.decoder.decoder.weight"] = _uniform(g, (A, 1, P), bound)
    sd["actor.decoder.decoder.bias"] = _uniform(g, (A,), bound)
    sd["actor.log_std"] = torch.full((A,), -0.001, dtype=torch.float32)
    if c.trained:
        sd["actor.log_std"] = sd["actor.log_std"] + _uniform(g, (A,), 0.1)
    return sd


def make_obs(c: Case) -> torch.Tensor:
    g = torch.Generator().manual_seed(2000 + c.seed)
    x = torch.randn(c.batch, c.obs_dim, generator=g, dtype=torch.float32)
    if c.obs_kind == "stress":
        # cover the encoder's +-5 receptive range, the env's +-100 clip
        # (AMP legged_robot_config.py:166) and exact receptive-field centres
        x = x * 3.0
        x[0::7] = x[0::7].clamp(-5, 5).round()
        x[1::7] = 100.0
        x[2::7] = -100.0
        x[3::7] = 0.0
    return x


def make_gmu(c: Case) -> torch.Tensor:
    """Upstream gradient dL/dmu used for the backward fixtures."""
    g = torch.Generator().manual_seed(3000 + c.seed)
    return torch.randn(c.batch, c.act_dim, generator=g, dtype=torch.float32)


# --------------------------------------------------------------------------- PPO update case
@dataclass(frozen=True)
class PPOCase:
    name: str = "ppo_cassie_tiny"
    variant: str = "cassie"
    obs_dim: int = 6
    hidden: tuple = (32, 24)
    critic_hidden: tuple = (16, 16)
    act_dim: int = 3
    num_envs: int = 8
    num_steps: int = 4
    epochs: int = 2
    seed: int = 21
    gamma: float = 0.99
    lam: float = 0.95
    lr: float = 1e-3
    entropy_coef: float = 0.01


PPO_CASE = PPOCase()


def make_ppo_state_dict(pc: PPOCase) -> Dict[str, torch.Tensor]:
    c = Case("ppo_actor", pc.variant, pc.obs_dim, pc.hidden, pc.act_dim, 1, seed=pc.seed)
    sd = make_state_dict(c)
    g = torch.Generator().manual_seed(4000 + pc.seed)
    dims = [pc.obs_dim] + list(pc.critic_hidden) + [1]
    for i in range(len(dims) - 1):
        bound = 1.0 / math.sqrt(dims[i])
        sd[f"critic.{2 * i}.weight"] = _uniform(g, (dims[i + 1], dims[i]), bound)
        sd[f"critic.{2 * i}.bias"] = _uniform(g, (dims[i + 1],), bound)
    sd["std"] = torch.ones(pc.act_dim)
    return sd


def make_ppo_rollout(pc: PPOCase):
    """Seeded rollout inputs that do not depend on the policy: obs, rewards, dones, action noise, last obs."""
    g = torch.Generator().manual_seed(5000 + pc.seed)
    S, N = pc.num_steps, pc.num_envs
    return {
        "obs": torch.randn(S, N, pc.obs_dim, generator=g),
        "rewards": torch.randn(S, N, generator=g),
        "dones": (torch.rand(S, N, generator=g) < 0.15).to(torch.uint8),
        "noise": torch.randn(S, N, pc.act_dim, generator=g),
        "last_obs": torch.randn(N, pc.obs_dim, generator=g),
    }


# --------------------------------------------------------------------------- RMA PPO update case
@dataclass(frozen=True)
class RMACase:
    """RMA fork's PPO.update (RMA-rl/algorithms/ppo.py:134-221): PPO step + adaptation-module regression substeps."""
    name: str = "rma_ppo_tiny"
    num_obs: int = 7
    num_priv: int = 5
    num_hist: int = 14
    latent: int = 4
    hidden: tuple = (32, 24)
    critic_hidden: tuple = (16, 16)
    enc_hidden: tuple = (12, 8)
    adapt_hidden: tuple = (12, 8)
    act_dim: int = 3
    num_envs: int = 8
    num_steps: int = 4
    epochs: int = 2
    substeps: int = 2
    seed: int = 23
    gamma: float = 0.99
    lam: float = 0.95
    lr: float = 1e-3
    entropy_coef: float = 0.01


RMA_CASE = RMACase()


def make_rma_state_dict(rc: RMACase) -> Dict[str, torch.Tensor]:
    c = Case("rma_actor", "rma", rc.num_obs + rc.latent, rc.hidden, rc.act_dim, 1, seed=rc.seed)
    sd = make_state_dict(c)
    g = torch.Generator().manual_seed(7000 + rc.seed)

    def mlp(prefix, dims):
        for i in range(len(dims) - 1):
            bound = 1.0 / math.sqrt(dims[i])
            sd[f"{prefix}.{2 * i}.weight"] = _uniform(g, (dims[i + 1], dims[i]), bound)
            sd[f"{prefix}.{2 * i}.bias"] = _uniform(g, (dims[i + 1],), bound)

    mlp("critic", [rc.num_obs + rc.latent] + list(rc.critic_hidden) + [1])
    mlp("env_factor_encoder", [rc.num_priv] + list(rc.enc_hidden) + [rc.latent])
    mlp("adaptation_module", [rc.num_hist] + list(rc.adapt_hidden) + [rc.latent])
    for k in [k for k in sd if k.startswith("env_factor_encoder.")]:       # the reference registers the module twice
        sd["encoder." + k[len("env_factor_encoder."):]] = sd[k]
    sd["std"] = torch.ones(rc.act_dim)
    return sd


def make_rma_rollout(rc: RMACase):
    g = torch.Generator().manual_seed(8000 + rc.seed)
    S, N = rc.num_steps, rc.num_envs
    return {
        "obs": torch.randn(S, N, rc.num_obs, generator=g),
        "priv": torch.randn(S, N, rc.num_priv, generator=g),
        "hist": torch.randn(S, N, rc.num_hist, generator=g),
        "rewards": torch.randn(S, N, generator=g),
        "dones": (torch.rand(S, N, generator=g) < 0.15).to(torch.uint8),
        "noise": torch.randn(S, N, rc.act_dim, generator=g),
        "last_obs": torch.randn(N, rc.num_obs, generator=g),
        "last_priv": torch.randn(N, rc.num_priv, generator=g),
    }


# --------------------------------------------------------------------------- AMP PPO update case
AMP_CASE = PPOCase(name="amp_ppo_tiny", variant="amp", seed=22)
AMP_OBS_DIM = 5
AMP_DISC_HIDDEN = (32, 16)
AMP_REWARD_COEF = 2.0
AMP_MIN_STD = 1.2            # above init_noise_std = 1.0, so the clamp is observable on ActorCritic.std
AMP_POOL = 64


def make_amp_inputs(pc: PPOCase):
    """Seeded AMP observations of the rollout (state at t, state at t+1), the expert pool and the discriminator."""
    g = torch.Generator().manual_seed(6000 + pc.seed)
    S, N, D = pc.num_steps, pc.num_envs, AMP_OBS_DIM
    out = {
        "amp_obs": torch.randn(S + 1, N, D, generator=g) * 1.5 + 0.3,
        "expert_state": torch.randn(AMP_POOL, D, generator=g) * 0.7 - 0.2,
        "expert_next": torch.randn(AMP_POOL, D, generator=g) * 0.7 - 0.1,
    }
    dims = [2 * D] + list(AMP_DISC_HIDDEN)
    sd = {}
    for i in range(len(dims) - 1):
        b = 1.0 / math.sqrt(dims[i])
        sd[f"trunk.{2 * i}.weight"] = _uniform(g, (dims[i + 1], dims[i]), b)
        sd[f"trunk.{2 * i}.bias"] = _uniform(g, (dims[i + 1],), b)
    b = 1.0 / math.sqrt(dims[-1])
    sd["amp_linear.weight"] = _uniform(g, (1, dims[-1]), b)
    sd["amp_linear.bias"] = _uniform(g, (1,), b)
    out["disc_state_dict"] = sd
    return out


class ExpertPool:
    """Deterministic stand-in for the reference's AMPLoader (needs pybullet_utils + mocap files): same
    `feed_forward_generator(num_mini_batch, mini_batch_size)` surface, rows picked by a fixed rule."""

    def __init__(self, state, next_state):
        self.state, self.next_state = state, next_state

    def feed_forward_generator(self, num_mini_batch, mini_batch_size):
        n = self.state.shape[0]
        for k in range(num_mini_batch):
            idx = (torch.arange(mini_batch_size, device=self.state.device) * 5 + 11 * k) % n
            yield self.state[idx], self.next_state[idx]


def deterministic_replay_generator(buf):
    """Replacement for ReplayBuffer.feed_forward_generator (random in both implementations) with a fixed index rule,
    installed on the reference's buffer by make_golden_amp.py and on ours by the tests."""

    def gen(num_mini_batch, mini_batch_size):
        for k in range(num_mini_batch):
            idx = (torch.arange(mini_batch_size, device=buf.states.device) * 3 + 7 * k) % buf.num_samples
            yield buf.states[idx], buf.next_states[idx]

    return gen
