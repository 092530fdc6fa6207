"""ActorCritic — drop-in for rsl_rl.modules.ActorCritic of the four reference forks.

Reference (cited from each fork's rsl_rl/modules/actor_critic.py):
  AMP      AMP/AMP_for_hardware-main/rsl_rl/rsl_rl/modules/actor_critic.py:323-430
  CASSIE   CASSIE/rsl_rl-master/rsl_rl/modules/actor_critic.py:322-426        (tanh on the decoder, :147)
  HUMANOID HUMANOID/pbrs-humanoid-dev/gpu_rl/rsl_rl/modules/actor_critic.py:323-451 (std^2 + 1e-5, :107)
  RMA      RMA/quadruped-robot-master/rl-robotics/rsl_rl/rsl_rl/modules/actor_critic.py:330-635
           (env_factor_encoder + adaptation_module; actor input = cat(obs, latent))

Same constructor signatures, methods, properties and state-dict keys; the actor is the fused sm_100a
PopSpikeActor, the critic stays a dense ELU MLP (it is not spiking in the reference either).
"""
from __future__ import annotations

import math
from typing import Sequence

import torch
import torch.nn as nn
from torch.distributions import Normal

from .actor import PopSpikeActor


def get_activation(act_name: str):
    table = {"elu": nn.ELU, "selu": nn.SELU, "relu": nn.ReLU, "crelu": nn.ReLU, "lrelu": nn.LeakyReLU,
             "tanh": nn.Tanh, "sigmoid": nn.Sigmoid}
    if act_name not in table:
        print("invalid activation function!")
        return None
    return table[act_name]()


def _mlp(in_dim: int, hidden: Sequence[int], out_dim: int, activation) -> nn.Sequential:
    """Linear -> act -> ... -> Linear(out_dim); module indices 0,2,4,... like the reference's critic (:359-368)."""
    layers = [nn.Linear(in_dim, hidden[0]), activation]
    for i in range(len(hidden)):
        if i == len(hidden) - 1:
            layers.append(nn.Linear(hidden[i], out_dim))
        else:
            layers.append(nn.Linear(hidden[i], hidden[i + 1]))
            layers.append(activation)
    return nn.Sequential(*layers)


class _SpikingActorCritic(nn.Module):
    is_recurrent = False
    # per-fork switches
    ENC_EPS = 0.0
    DEC_TANH = False

    def _build(self, actor_in: int, critic_in: int, num_actions: int, actor_hidden_dims, critic_hidden_dims,
               activation, init_noise_std: float, fixed_std: bool = False, spike_ts: int = 5, en_pop: int = 10,
               de_pop: int = 10, mean_range=(-5, 5), enc_std: float = math.sqrt(0.15), fwd_planes: int = 3):
        act = get_activation(activation)
        self.actor = PopSpikeActor(actor_in, num_actions, en_pop, de_pop, list(actor_hidden_dims), mean_range,
                                   enc_std, spike_ts=spike_ts, device="cuda", enc_eps=self.ENC_EPS,
                                   dec_tanh=self.DEC_TANH, fwd_planes=fwd_planes)
        self.critic = _mlp(critic_in, list(critic_hidden_dims), 1, act)
        # `std` is dead in SNN mode (the actor owns log_std) but runners log / clamp it, so it must exist
        self.fixed_std = fixed_std
        std = init_noise_std * torch.ones(num_actions)
        if fixed_std:
            self.std = std.clone()
        else:
            self.std = nn.Parameter(std)
        self.distribution = None

    @staticmethod
    def _warn_unexpected(kwargs):
        if kwargs:
            print("ActorCritic.__init__ got unexpected arguments, which will be ignored: "
                  + str([key for key in kwargs.keys()]))

    def reset(self, dones=None):
        pass

    def forward(self):
        raise NotImplementedError

    @property
    def action_mean(self):
        return self.distribution.mean

    @property
    def action_std(self):
        return self.distribution.stddev

    @property
    def entropy(self):
        return self.distribution.entropy().sum(dim=-1)

    def _set_distribution(self, actor_input):
        mean, std = self.actor(actor_input)
        # validate_args=False: the argument check is a host synchronisation (`constraint.check(...).all()`), which the
        # reference tries to switch off at :378 (`Normal.set_default_validate_args = False` assigns instead of calling)
        self.distribution = Normal(mean, mean * 0. + std, validate_args=False)

    def update_distribution(self, observations):
        self._set_distribution(observations)

    def act(self, observations, **kwargs):
        self.update_distribution(observations)
        return self.distribution.sample()

    def get_actions_log_prob(self, actions):
        return self.distribution.log_prob(actions).sum(dim=-1)

    def act_inference(self, observations):
        actions_mean, _ = self.actor(observations)
        return actions_mean

    def evaluate(self, critic_observations, **kwargs):
        return self.critic(critic_observations)


class ActorCritic(_SpikingActorCritic):
    """AMP fork signature (AMP .../actor_critic.py:325-335)."""

    def __init__(self, num_actor_obs, num_critic_obs, num_actions, actor_hidden_dims=[256, 256, 256],
                 critic_hidden_dims=[256, 256, 256], activation='elu', init_noise_std=1.0, fixed_std=False,
                 **kwargs):
        snn_kw = {k: kwargs.pop(k) for k in ("spike_ts", "en_pop", "de_pop", "fwd_planes") if k in kwargs}
        self._warn_unexpected(kwargs)
        super().__init__()
        self._build(num_actor_obs, num_critic_obs, num_actions, actor_hidden_dims, critic_hidden_dims, activation,
                    init_noise_std, fixed_std=fixed_std, **snn_kw)


ActorCriticAMP = ActorCritic


class ActorCriticCassie(_SpikingActorCritic):
    """CASSIE fork (CASSIE/rsl_rl-master/.../actor_critic.py:324-331): tanh on mu, Normal(mean, std) broadcast."""
    DEC_TANH = True

    def __init__(self, num_actor_obs, num_critic_obs, num_actions, actor_hidden_dims=[256, 256, 256],
                 critic_hidden_dims=[256, 256, 256], activation='elu', init_noise_std=1.0, **kwargs):
        snn_kw = {k: kwargs.pop(k) for k in ("spike_ts", "en_pop", "de_pop", "fwd_planes") if k in kwargs}
        self._warn_unexpected(kwargs)
        super().__init__()
        self._build(num_actor_obs, num_critic_obs, num_actions, actor_hidden_dims, critic_hidden_dims, activation,
                    init_noise_std, **snn_kw)

    def _set_distribution(self, actor_input):
        mean, std = self.actor(actor_input)
        self.distribution = Normal(mean, std, validate_args=False)        # CASSIE :410 — std broadcasts


class ActorCriticHumanoid(_SpikingActorCritic):
    """HUMANOID fork (pbrs-humanoid .../actor_critic.py:325-334): encoder uses std^2 + 1e-5."""
    ENC_EPS = 1e-5

    def __init__(self, num_actor_obs, num_critic_obs, num_actions, actor_hidden_dims=[256, 256, 256],
                 critic_hidden_dims=[256, 256, 256], activation='elu', init_noise_std=1.0,
                 custom_actor_args={}, custom_critic_args={}, **kwargs):
        snn_kw = {k: kwargs.pop(k) for k in ("spike_ts", "en_pop", "de_pop", "fwd_planes") if k in kwargs}
        self._warn_unexpected(kwargs)
        if custom_actor_args or custom_critic_args:
            raise NotImplementedError("custom_actor_args / custom_critic_args (ROM residual networks) are outside "
                                      "the spiking hot path (SURVEY.md §2.2)")
        super().__init__()
        self._build(num_actor_obs, num_critic_obs, num_actions, actor_hidden_dims, critic_hidden_dims, activation,
                    init_noise_std, **snn_kw)

    def _set_distribution(self, actor_input):
        mean, std = self.actor(actor_input)
        self.distribution = Normal(mean, std, validate_args=False)        # HUMANOID :430


class ActorCriticRMA(_SpikingActorCritic):
    """RMA fork (RMA .../actor_critic.py:372-635): privileged-obs encoder + adaptation module; the actor's input
    is cat(obs, latent), so d loss / d actor-input flows back into env_factor_encoder."""

    def __init__(self, num_obs, num_privileged_obs, num_obs_history, num_actions, actor_hidden_dims=[256, 256, 256],
                 critic_hidden_dims=[256, 256, 256], encoder_hidden_dims=[256, 128], adaptation_hidden_dims=[256, 32],
                 encoder_input_dims=50, encoder_latent_dims=18, activation='elu', init_noise_std=1.0, **kwargs):
        snn_kw = {k: kwargs.pop(k) for k in ("spike_ts", "en_pop", "de_pop", "fwd_planes") if k in kwargs}
        self._warn_unexpected(kwargs)
        super().__init__()
        act = get_activation(activation)
        self.env_factor_encoder = _mlp(num_privileged_obs, list(encoder_hidden_dims), encoder_latent_dims, act)
        self.add_module("encoder", self.env_factor_encoder)          # the reference registers it twice (:404)
        self.adaptation_module = _mlp(num_obs_history, list(adaptation_hidden_dims), encoder_latent_dims, act)
        latent_dim = int(encoder_latent_dims)
        self._build(num_obs + latent_dim, num_obs + latent_dim, num_actions, actor_hidden_dims, critic_hidden_dims,
                    activation, init_noise_std, **snn_kw)

    def update_distribution(self, observations, privileged_observations):
        latent = self.env_factor_encoder(privileged_observations)
        self._set_distribution(torch.cat((observations, latent), dim=-1))

    def act(self, observations, privileged_observations, **kwargs):
        self.update_distribution(observations, privileged_observations)
        return self.distribution.sample()

    def act_expert(self, observations, privileged_observations, policy_info={}):
        latent = self.env_factor_encoder(privileged_observations)
        actions_mean, _ = self.actor(torch.cat((observations, latent), dim=-1))
        policy_info["latents"] = latent.detach().cpu().numpy()
        return actions_mean

    def act_inference(self, observations, observation_history, privileged_observations=None, policy_info={}):
        if privileged_observations is not None:
            latent = self.env_factor_encoder(privileged_observations)
            policy_info["gt_latents"] = latent.detach().cpu().numpy()
        latent = self.adaptation_module(observation_history)
        actions_mean, _ = self.actor(torch.cat((observations, latent), dim=-1))
        policy_info["latents"] = latent.detach().cpu().numpy()
        return actions_mean

    def evaluate(self, critic_observations, privileged_observations, **kwargs):
        latent = self.env_factor_encoder(privileged_observations)
        return self.critic(torch.cat((critic_observations, latent), dim=-1))


VARIANTS = {"amp": ActorCritic, "cassie": ActorCriticCassie, "humanoid": ActorCriticHumanoid, "rma": ActorCriticRMA}
