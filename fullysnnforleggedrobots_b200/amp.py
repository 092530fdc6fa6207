"""AMP fork's PPO — same surface as rsl_rl.algorithms.AMPPPO and its helpers (reference:
AMP/AMP_for_hardware-main/rsl_rl/rsl_rl/algorithms/amp_ppo.py:38-274, algorithms/amp_discriminator.py:9-72,
storage/replay_buffer.py:5-50, utils/utils.py:79-131).

What runs where:
  * the spiking actor-critic part of every minibatch step (actor forward + BPTT, critic, PPO head, adaptive-KL
    learning rate, clip_grad_norm_ over the actor-critic parameters only, Adam) is `PPO`'s fused CUDA-graph step
    on the sm_100a kernels — unchanged;
  * the discriminator (dense ReLU MLP, least-squares GAN loss, gradient penalty = a double backward) stays torch
    (not spiking, not on the kernel path).  The reference keeps ONE Adam with three parameter
    groups (actor-critic; trunk, weight_decay 10e-4; head, weight_decay 10e-2) sharing the adaptive learning rate.
    Adam is per-parameter, so two optimizers with the same learning rate are the same update: the discriminator's
    Adam reads the learning rate from the device scalar `snn_lr_adapt` maintains (no host round trip).
  * `Normalizer` keeps its running moments on the device (float64); the reference converts every minibatch to
    numpy on the host (`amp_ppo.py:254-256`, two device->host copies + syncs per step).

Quirks of the reference that are kept on purpose (parity): the normaliser is updated with the already NORMALISED
policy / expert states (`amp_ppo.py:219-226,254-256` rebinds the names before the update), the gradient penalty
uses the raw expert sample (`:236-237`), and `min_std` clamps `ActorCritic.std`, which is dead in SNN mode
(the actor owns `log_std`).
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch
import torch.nn as nn
from torch import autograd

from .ppo import PPO
from .storage import RolloutStorage


class AMPDiscriminator(nn.Module):
    """amp_discriminator.py:9-72 (state-dict keys `trunk.{0,2,..}`, `amp_linear`)."""

    def __init__(self, input_dim, amp_reward_coef, hidden_layer_sizes, device, task_reward_lerp=0.0):
        super().__init__()
        self.device = device
        self.input_dim = input_dim
        self.amp_reward_coef = amp_reward_coef
        layers, d = [], input_dim
        for h in hidden_layer_sizes:
            layers += [nn.Linear(d, h), nn.ReLU()]
            d = h
        self.trunk = nn.Sequential(*layers).to(device)
        self.amp_linear = nn.Linear(hidden_layer_sizes[-1], 1).to(device)
        self.trunk.train()
        self.amp_linear.train()
        self.task_reward_lerp = task_reward_lerp

    def forward(self, x):
        return self.amp_linear(self.trunk(x))

    def compute_grad_pen(self, expert_state, expert_next_state, lambda_=10):
        expert_data = torch.cat([expert_state, expert_next_state], dim=-1)
        expert_data.requires_grad = True
        disc = self.amp_linear(self.trunk(expert_data))
        grad = autograd.grad(outputs=disc, inputs=expert_data, grad_outputs=torch.ones_like(disc), create_graph=True,
                             retain_graph=True, only_inputs=True)[0]
        return lambda_ * grad.norm(2, dim=1).pow(2).mean()

    def predict_amp_reward(self, state, next_state, task_reward, normalizer=None):
        with torch.no_grad():
            self.eval()
            if normalizer is not None:
                state = normalizer.normalize_torch(state, self.device)
                next_state = normalizer.normalize_torch(next_state, self.device)
            d = self.amp_linear(self.trunk(torch.cat([state, next_state], dim=-1)))
            reward = self.amp_reward_coef * torch.clamp(1 - (1 / 4) * torch.square(d - 1), min=0)
            if self.task_reward_lerp > 0:
                reward = self._lerp_reward(reward, task_reward.unsqueeze(-1))
            self.train()
        return reward.squeeze(), d

    def _lerp_reward(self, disc_r, task_r):
        return (1.0 - self.task_reward_lerp) * disc_r + self.task_reward_lerp * task_r


class ReplayBuffer:
    """Fixed-size ring of (state, next_state) AMP observations (replay_buffer.py:5-50).  Sampling draws the indices
    on the device (`torch.randint`; the reference uses `np.random.choice` on the host — same distribution, with
    replacement, different stream)."""

    def __init__(self, obs_dim, buffer_size, device):
        self.states = torch.zeros(buffer_size, obs_dim, device=device)
        self.next_states = torch.zeros(buffer_size, obs_dim, device=device)
        self.buffer_size = buffer_size
        self.device = device
        self.step = 0
        self.num_samples = 0

    def insert(self, states, next_states):
        """Ring write of n rows starting at the cursor: one scatter per field (row i -> (cursor + i) mod size), which
        is the reference's two-slice wrap-around without the branch.  More rows than the ring holds: the last ones
        win, as a sequential write would leave them."""
        n = states.shape[0]
        if n > self.buffer_size:
            states, next_states = states[n - self.buffer_size:], next_states[n - self.buffer_size:]
            self.step = (self.step + n - self.buffer_size) % self.buffer_size
            self.num_samples = self.buffer_size
            n = self.buffer_size
        rows = (torch.arange(n, device=self.states.device) + self.step) % self.buffer_size
        self.states.index_copy_(0, rows, states.to(self.states.dtype))
        self.next_states.index_copy_(0, rows, next_states.to(self.states.dtype))
        self.num_samples = min(self.buffer_size, max(self.step + n, self.num_samples))
        self.step = (self.step + n) % self.buffer_size

    def feed_forward_generator(self, num_mini_batch, mini_batch_size):
        for _ in range(num_mini_batch):
            idx = torch.randint(0, self.num_samples, (mini_batch_size,), device=self.states.device)
            yield self.states.index_select(0, idx), self.next_states.index_select(0, idx)


class Normalizer:
    """Running mean / variance of the AMP observations (utils.py:79-131), parallel-variance update, kept on the
    device in float64 and updated IN PLACE (mean, var and count are device tensors with stable addresses, so the
    discriminator step that uses them can be replayed as a CUDA graph).  `update` takes a torch tensor (any device)
    or a numpy array; nothing here synchronises (`float(normalizer.count)` does, when a caller asks)."""

    def __init__(self, input_dim, epsilon=1e-4, clip_obs=10.0, device="cpu"):
        self.device = torch.device(device)
        self.mean = torch.zeros(input_dim, dtype=torch.float64, device=self.device)
        self.var = torch.ones(input_dim, dtype=torch.float64, device=self.device)
        self.count = torch.full((), 1e-4, dtype=torch.float64, device=self.device)   # RunningMeanStd's epsilon (:80,89)
        self.epsilon = epsilon
        self.clip_obs = clip_obs

    def to(self, device):
        self.device = torch.device(device)
        self.mean, self.var, self.count = (t.to(self.device) for t in (self.mean, self.var, self.count))
        return self

    @torch.no_grad()
    def update(self, arr):
        x = torch.as_tensor(arr).to(device=self.device, dtype=torch.float64)
        self.update_from_moments(x.mean(dim=0), x.var(dim=0, unbiased=False), x.shape[0])

    @torch.no_grad()
    def update_from_moments(self, batch_mean, batch_var, batch_count):
        delta = batch_mean - self.mean
        tot = self.count + batch_count
        m_2 = self.var * self.count + batch_var * batch_count + delta.square() * (self.count * batch_count / tot)
        self.mean.add_(delta * (batch_count / tot))
        self.var.copy_(m_2 / tot)
        self.count.copy_(tot)

    def normalize(self, input):
        mean, var = self.mean.cpu().numpy(), self.var.cpu().numpy()
        return np.clip((input - mean) / np.sqrt(var + self.epsilon), -self.clip_obs, self.clip_obs)

    def normalize_torch(self, input, device=None):
        dev = input.device if device is None else device
        mean = self.mean.to(device=dev, dtype=torch.float32)
        std = torch.sqrt((self.var + self.epsilon).to(device=dev, dtype=torch.float32))
        return torch.clamp((input - mean) / std, -self.clip_obs, self.clip_obs)


class AMPPPO(PPO):
    """AMPPPO (amp_ppo.py:38-274): PPO on the spiking actor-critic + discriminator loss + AMP replay buffer.
    `amp_data` is any object with `feed_forward_generator(num_mini_batch, mini_batch_size)` yielding
    `(state, next_state)` expert batches (the reference's AMPLoader needs pybullet_utils and mocap files: out of
    scope).  Returns the reference's 6-tuple from `update()`."""

    def __init__(self, actor_critic, discriminator, amp_data, amp_normalizer, num_learning_epochs=1,
                 num_mini_batches=1, clip_param=0.2, gamma=0.998, lam=0.95, value_loss_coef=1.0, entropy_coef=0.0,
                 learning_rate=1e-3, max_grad_norm=1.0, use_clipped_value_loss=True, schedule="fixed", desired_kl=0.01,
                 device='cpu', amp_replay_buffer_size=100000, min_std=None, disc_matmul="tf32", **kwargs):
        super().__init__(actor_critic, num_learning_epochs=num_learning_epochs, num_mini_batches=num_mini_batches,
                         clip_param=clip_param, gamma=gamma, lam=lam, value_loss_coef=value_loss_coef,
                         entropy_coef=entropy_coef, learning_rate=learning_rate, max_grad_norm=max_grad_norm,
                         use_clipped_value_loss=use_clipped_value_loss, schedule=schedule, desired_kl=desired_kl,
                         device=device, **kwargs)
        self.min_std = min_std
        # Discriminator GEMMs (60 -> 1024 -> 512 -> 1 on 2 x 24576 rows, plus the double backward of the gradient
        # penalty) are 90 % of an AMP update when cuBLAS runs them as fp32 SIMT sgemm (measured: 205 ms vs 14 ms for the
        # spiking PPO part).  "tf32" lets cuBLAS use the tensor cores for them — which is what the reference's pinned
        # torch 1.10 does by default on Ampere and later (allow_tf32 was on until torch 1.12); "fp32" = IEEE sgemm.
        if disc_matmul not in ("tf32", "fp32"):
            raise ValueError("disc_matmul must be 'tf32' or 'fp32'")
        self.disc_matmul = disc_matmul
        self.discriminator = discriminator
        self.discriminator.to(self.device)
        if self.dp is not None:
            self.dp.broadcast_parameters(self.discriminator)
        self.amp_transition = RolloutStorage.Transition()
        self.amp_storage = ReplayBuffer(discriminator.input_dim // 2, amp_replay_buffer_size, device)
        self.amp_data = amp_data
        self.amp_normalizer = amp_normalizer
        if amp_normalizer is not None and hasattr(amp_normalizer, "to"):
            amp_normalizer.to(self.device)
        # the discriminator's two parameter groups of the reference Adam (:87-92); lr = the device scalar that
        # snn_lr_adapt updates inside the fused step (0-dim view of FlatAdam.lr_dev)
        self.disc_optimizer = torch.optim.Adam(
            [{'params': self.discriminator.trunk.parameters(), 'weight_decay': 10e-4, 'name': 'amp_trunk'},
             {'params': self.discriminator.amp_linear.parameters(), 'weight_decay': 10e-2, 'name': 'amp_head'}],
            lr=self.optimizer.lr_dev[0], capturable=True, foreach=False)
        self._amp_stats = torch.zeros(4, dtype=torch.float32, device=self.optimizer.flat_params.device)
        # the discriminator step (normalise, two forwards, gradient penalty, backward, Adam, normaliser update:
        # ~150 small launches, host-bound when launched eagerly) is replayed as ONE CUDA graph from its third call on
        self.disc_cuda_graph = self.use_cuda_graph
        self._disc_state = None

    def test_mode(self):
        self.actor_critic.eval()           # the reference calls the non-existent `.test()` here (:111)

    def act(self, obs, critic_obs, amp_obs):
        actions = super().act(obs.detach(), critic_obs.detach())
        self.amp_transition.observations = amp_obs
        return actions

    def process_env_step(self, rewards, dones, infos, amp_obs):
        self.amp_storage.insert(self.amp_transition.observations, amp_obs)
        super().process_env_step(rewards, dones, infos)
        self.amp_transition.clear()

    # ------------------------------------------------------------------ discriminator half of a minibatch step
    def _disc_step_body(self, sample_amp_policy, sample_amp_expert):
        """Everything of a minibatch step that belongs to the discriminator; runs after the actor-critic's step (the
        two are independent apart from the learning rate, which that step has just adapted on the device)."""
        policy_state, expert_state = self._disc_backward_impl(sample_amp_policy, sample_amp_expert)
        self.disc_optimizer.step()
        if self.amp_normalizer is not None:
            self.amp_normalizer.update(policy_state)
            self.amp_normalizer.update(expert_state)

    def _disc_graphable(self):
        return (self.disc_cuda_graph and self.dp is None and self._amp_stats.is_cuda
                and (self.amp_normalizer is None or isinstance(self.amp_normalizer, Normalizer)))

    def _disc_step(self, sample_amp_policy, sample_amp_expert):
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = self.disc_matmul == "tf32"
        try:
            if not self._disc_graphable():
                return self._disc_step_body(sample_amp_policy, sample_amp_expert)
            tensors = (*sample_amp_policy, *sample_amp_expert)
            key = tuple((tuple(t.shape), t.dtype) for t in tensors)
            st = self._disc_state
            if st is None or st["key"] != key:
                st = self._disc_state = {"key": key, "calls": 0, "graph": None,
                                         "inputs": [torch.empty_like(t) for t in tensors]}
            for dst, src in zip(st["inputs"], tensors):
                dst.copy_(src)
            ins = st["inputs"]
            st["calls"] += 1
            if st["graph"] is None:
                if st["calls"] <= 2:      # real steps as warm-up (lazy Adam state, cuBLAS workspaces), on a side stream
                    side = torch.cuda.Stream()
                    side.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(side):
                        self._disc_step_body((ins[0], ins[1]), (ins[2], ins[3]))
                    torch.cuda.current_stream().wait_stream(side)
                    return
                try:
                    torch.cuda.synchronize()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):              # capture only records: the step itself runs at replay()
                        self._disc_step_body((ins[0], ins[1]), (ins[2], ins[3]))
                    st["graph"] = g
                except Exception as e:  # noqa: BLE001
                    print(f"[AMPPPO] CUDA graph capture of the discriminator step failed ({e}); running it eagerly")
                    self.disc_cuda_graph = False
                    return self._disc_step_body(sample_amp_policy, sample_amp_expert)
            st["graph"].replay()
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev

    def _disc_backward_impl(self, sample_amp_policy, sample_amp_expert):
        policy_state, policy_next_state = sample_amp_policy
        expert_state, expert_next_state = sample_amp_expert
        norm = self.amp_normalizer
        if norm is not None:
            with torch.no_grad():
                policy_state = norm.normalize_torch(policy_state, self.device)
                policy_next_state = norm.normalize_torch(policy_next_state, self.device)
                expert_state = norm.normalize_torch(expert_state, self.device)
                expert_next_state = norm.normalize_torch(expert_next_state, self.device)
        disc = self.discriminator
        policy_d = disc(torch.cat([policy_state, policy_next_state], dim=-1))
        expert_d = disc(torch.cat([expert_state, expert_next_state], dim=-1))
        expert_loss = torch.nn.functional.mse_loss(expert_d, torch.ones_like(expert_d))
        policy_loss = torch.nn.functional.mse_loss(policy_d, -torch.ones_like(policy_d))
        amp_loss = 0.5 * (expert_loss + policy_loss)
        grad_pen_loss = disc.compute_grad_pen(*sample_amp_expert, lambda_=10)
        self.disc_optimizer.zero_grad(set_to_none=True)
        (amp_loss + grad_pen_loss).backward()
        if self.dp is not None:
            self.dp.allreduce_gradients(disc)
        with torch.no_grad():
            self._amp_stats += torch.stack([amp_loss.detach(), grad_pen_loss.detach(), policy_d.mean(),
                                            expert_d.mean()])
        return policy_state, expert_state

    def update(self):
        self._stats.zero_()
        self._amp_stats.zero_()
        self.optimizer.set_lr(self._lr)
        ac = self.actor_critic
        num_updates = self.num_learning_epochs * self.num_mini_batches
        mbs = self.storage.num_envs * self.storage.num_transitions_per_env // self.num_mini_batches
        amp_policy_generator = self.amp_storage.feed_forward_generator(num_updates, mbs)
        amp_expert_generator = self.amp_data.feed_forward_generator(num_updates, mbs)
        fast = self._fast_supported()
        if fast:
            f, mb = self.storage.permute_once(self.num_mini_batches)
            if self._fast_state is None or self._fast_state["mb"] != mb:
                self._fast_setup(mb)
            self.refresh_packed()
            self.optimizer.zero_grad()     # once per update (PPO._step_fast: the fused kernels overwrite their gradients)
            steps = ((lambda i=i: self._run_fast(f, i)) for _ in range(self.num_learning_epochs)
                     for i in range(self.num_mini_batches))
        else:
            steps = ((lambda b=b: self._step_generic(b))
                     for b in self.storage.mini_batch_generator(self.num_mini_batches, self.num_learning_epochs))
        for ac_step, sample_amp_policy, sample_amp_expert in zip(steps, amp_policy_generator, amp_expert_generator):
            # spiking actor-critic: forward, BPTT, adaptive lr -> lr_dev, clip over ITS parameters only (:248), Adam
            ac_step()
            # discriminator: losses, gradients, the same Adam step for its two groups at the learning rate the
            # actor-critic step just set, normaliser update
            self._disc_step(sample_amp_policy, sample_amp_expert)
            if not ac.fixed_std and self.min_std is not None:
                ac.std.data.clamp_(min=self.min_std)
        stats = self._stats.tolist()                       # the host syncs of the update: here ...
        amp = self._amp_stats.tolist()
        self._lr = self.optimizer.sync_lr_from_device()    # ... and here
        self._params_epoch += 1
        self.storage.clear()
        return (stats[5] / num_updates, stats[4] / num_updates, amp[0] / num_updates, amp[1] / num_updates,
                amp[2] / num_updates, amp[3] / num_updates)
