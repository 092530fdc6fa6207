"""RolloutStorage — same public surface as rsl_rl.storage.RolloutStorage (reference:
AMP/AMP_for_hardware-main/rsl_rl/rsl_rl/storage/rollout_storage.py:36-184; the CASSIE / HUMANOID copies are
identical, the RMA copy adds observation histories), laid out for the B200 update kernels:

  * buffers stay [steps, envs, feature] fp32 (uint8 dones), contiguous, so every rollout step is one coalesced
    row block and GAE is a single reverse-scan kernel (snn_gae) instead of a Python loop of ~8 kernels/step;
  * minibatching permutes ONCE per update: the reference re-gathers the same random rows for each of the
    num_epochs passes (the permutation is drawn once, :151); here each field is gathered one time into a
    minibatch-contiguous buffer and every (epoch, minibatch) is a zero-copy row slice of it.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from . import engine as _engine


class RolloutStorage:
    class Transition:
        def __init__(self):
            self.observations = None
            self.critic_observations = None
            self.actions = None
            self.rewards = None
            self.dones = None
            self.values = None
            self.actions_log_prob = None
            self.action_mean = None
            self.action_sigma = None
            self.hidden_states = None
            self.observation_histories = None      # RMA only
            self.privileged_observations = None    # RMA only

        def clear(self):
            self.__init__()

    def __init__(self, num_envs, num_transitions_per_env, obs_shape, privileged_obs_shape, actions_shape,
                 device='cpu', obs_history_shape=None):
        self.device = device
        self.obs_shape = obs_shape
        self.privileged_obs_shape = privileged_obs_shape
        self.actions_shape = actions_shape
        S, N = num_transitions_per_env, num_envs
        z = lambda *shape, dtype=torch.float32: torch.zeros(S, N, *shape, device=device, dtype=dtype)
        self.observations = z(*obs_shape)
        self.privileged_observations = z(*privileged_obs_shape) if privileged_obs_shape[0] is not None else None
        self.observation_histories = z(*obs_history_shape) if obs_history_shape is not None else None
        self.rewards = z(1)
        self.actions = z(*actions_shape)
        self.dones = z(1, dtype=torch.uint8)
        self.actions_log_prob = z(1)
        self.values = z(1)
        self.returns = z(1)
        self.advantages = z(1)
        self.mu = z(*actions_shape)
        self.sigma = z(*actions_shape)
        self.num_transitions_per_env = S
        self.num_envs = N
        self.step = 0
        # data-parallel hook: maps local (sum, sum_sq, count) -> global, see dist.py
        self.advantage_stats_allreduce = None

    def add_transitions(self, transition: "RolloutStorage.Transition"):
        if self.step >= self.num_transitions_per_env:
            raise AssertionError("Rollout buffer overflow")
        s = self.step
        self.observations[s].copy_(transition.observations)
        if self.privileged_observations is not None:
            self.privileged_observations[s].copy_(transition.critic_observations)
        if self.observation_histories is not None and transition.observation_histories is not None:
            self.observation_histories[s].copy_(transition.observation_histories)
        self.actions[s].copy_(transition.actions)
        self.rewards[s].copy_(transition.rewards.view(-1, 1))
        self.dones[s].copy_(transition.dones.view(-1, 1))
        self.values[s].copy_(transition.values)
        self.actions_log_prob[s].copy_(transition.actions_log_prob.view(-1, 1))
        self.mu[s].copy_(transition.action_mean)
        self.sigma[s].copy_(transition.action_sigma)
        self.step += 1

    def clear(self):
        self.step = 0

    def compute_returns(self, last_values, gamma, lam):
        """GAE + advantage normalisation (reference :124-138) as one fused scan kernel + a 2-pass normaliser."""
        S, N = self.num_transitions_per_env, self.num_envs
        normalize = self.advantage_stats_allreduce is None
        ret, adv = _engine.gae(self.rewards.view(S, N), self.dones.view(S, N), self.values.view(S, N),
                               last_values.reshape(N).contiguous().float(), gamma, lam, normalize=normalize)
        self.returns = ret.view(S, N, 1)
        if not normalize:
            # data parallel: global mean / unbiased std over all ranks' samples
            from .dist import normalize_advantages_global
            adv = normalize_advantages_global(adv, self.advantage_stats_allreduce)
        self.advantages = adv.view(S, N, 1)

    def get_statistics(self):
        done = self.dones
        done[-1] = 1
        flat_dones = done.permute(1, 0, 2).reshape(-1, 1)
        done_indices = torch.cat((flat_dones.new_tensor([-1], dtype=torch.int64),
                                  flat_dones.nonzero(as_tuple=False)[:, 0]))
        trajectory_lengths = (done_indices[1:] - done_indices[:-1])
        return trajectory_lengths.float().mean(), self.rewards.mean()

    def _fields(self):
        obs = self.observations.flatten(0, 1)
        crit = self.privileged_observations.flatten(0, 1) if self.privileged_observations is not None else None
        return obs, crit

    def permute_once(self, num_mini_batches):
        """Draw the update's permutation (reference :151) and gather every field ONCE into persistent,
        minibatch-contiguous buffers (static addresses: the fused update replays CUDA graphs over them).
        Returns (dict of [num_mini_batches * mini_batch_size, feat] tensors, mini_batch_size)."""
        batch_size = self.num_envs * self.num_transitions_per_env
        mini_batch_size = batch_size // num_mini_batches
        n = num_mini_batches * mini_batch_size
        indices = torch.randperm(n, requires_grad=False, device=self.device)
        src = {"obs": self.observations, "critic_obs": self.privileged_observations, "actions": self.actions,
               "values": self.values, "returns": self.returns, "logp": self.actions_log_prob,
               "adv": self.advantages, "mu": self.mu, "sigma": self.sigma, "hist": self.observation_histories}
        if not hasattr(self, "_perm") or self._perm_n != n:
            self._perm = {k: torch.empty(n, *v.shape[2:], device=self.device, dtype=v.dtype)
                          for k, v in src.items() if v is not None}
            self._perm_n = n
        if self.observations.is_cuda:
            # all fields by ONE kernel (snn_gather_rows): the reference's nine `tensor[batch_idx]` gathers x 20
            keys = list(self._perm.keys())
            k = len(keys)
            srcs = (C.c_void_p * k)(*[src[x].data_ptr() for x in keys])
            dsts = (C.c_void_p * k)(*[self._perm[x].data_ptr() for x in keys])
            nbytes = (C.c_int32 * k)(*[self._perm[x][0].numel() * self._perm[x].element_size() for x in keys])
            dev = self.observations.device
            with torch.cuda.device(dev):
                _lib.check(_lib.lib().snn_gather_rows(k, srcs, dsts, nbytes, C.c_void_p(indices.data_ptr()), n,
                                                      C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)),
                           "snn_gather_rows")
        else:
            for k, buf in self._perm.items():
                torch.index_select(src[k].flatten(0, 1), 0, indices, out=buf)
        out = dict(self._perm)
        if "critic_obs" not in out:
            out["critic_obs"] = out["obs"]
        return out, mini_batch_size

    def mini_batch_generator(self, num_mini_batches, num_epochs=8):
        """Yields the reference's 11-tuple (:183-184); rows are permuted once, minibatches are contiguous slices."""
        batch_size = self.num_envs * self.num_transitions_per_env
        mini_batch_size = batch_size // num_mini_batches
        indices = torch.randperm(num_mini_batches * mini_batch_size, requires_grad=False, device=self.device)
        take = lambda t: t.flatten(0, 1).index_select(0, indices)
        obs = take(self.observations)
        crit = take(self.privileged_observations) if self.privileged_observations is not None else obs
        actions, values, returns = take(self.actions), take(self.values), take(self.returns)
        logp, adv, mu, sigma = take(self.actions_log_prob), take(self.advantages), take(self.mu), take(self.sigma)
        for _ in range(num_epochs):
            for i in range(num_mini_batches):
                sl = slice(i * mini_batch_size, (i + 1) * mini_batch_size)
                yield (obs[sl], crit[sl], actions[sl], values[sl], adv[sl], returns[sl], logp[sl], mu[sl],
                       sigma[sl], (None, None), None)


class RolloutStorageRMA(RolloutStorage):
    """RMA fork's storage (RMA/quadruped-robot-master/rl-robotics/rsl_rl/rsl_rl/storage/rollout_storage.py:38-269):
    privileged observations and observation histories are always stored, the critic sees the plain observations,
    and the generator yields the 13-tuple of :210-211."""

    def __init__(self, num_envs, num_transitions_per_env, obs_shape, privileged_obs_shape, obs_history_shape,
                 actions_shape, device='cpu'):
        super().__init__(num_envs, num_transitions_per_env, obs_shape, privileged_obs_shape, actions_shape, device,
                         obs_history_shape=obs_history_shape)

    def add_transitions(self, transition):
        # RMA keeps the privileged observations in their own field (reference :113-115)
        t = transition
        crit = t.critic_observations
        t.critic_observations = getattr(t, "privileged_observations", None)
        try:
            super().add_transitions(t)
        finally:
            t.critic_observations = crit

    def mini_batch_generator(self, num_mini_batches, num_epochs=8):
        batch_size = self.num_envs * self.num_transitions_per_env
        mini_batch_size = batch_size // num_mini_batches
        indices = torch.randperm(num_mini_batches * mini_batch_size, requires_grad=False, device=self.device)
        take = lambda t: t.flatten(0, 1).index_select(0, indices)
        obs = take(self.observations)
        priv = take(self.privileged_observations)
        hist = take(self.observation_histories)
        actions, values, returns = take(self.actions), take(self.values), take(self.returns)
        logp, adv, mu, sigma = take(self.actions_log_prob), take(self.advantages), take(self.mu), take(self.sigma)
        for _ in range(num_epochs):
            for i in range(num_mini_batches):
                sl = slice(i * mini_batch_size, (i + 1) * mini_batch_size)
                yield (obs[sl], obs[sl], priv[sl], hist[sl], actions[sl], values[sl], adv[sl], returns[sl], logp[sl],
                       mu[sl], sigma[sl], (None, None), None)
