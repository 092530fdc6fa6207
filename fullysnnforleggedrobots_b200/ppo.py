"""PPO — same surface as rsl_rl.algorithms.PPO (reference: AMP/AMP_for_hardware-main/rsl_rl/rsl_rl/algorithms/
ppo.py:38-187; CASSIE identical but default lr, HUMANOID adds Adam weight_decay=1e-5 at :70-72).

act / process_env_step / compute_returns / update keep the reference's semantics.  What changes underneath:
the actor forward/backward are the fused sm_100a kernels, GAE is one kernel, minibatches are contiguous
slices, the two per-minibatch `.item()` host syncs (:179-180) are deferred to one sync per update, and under
data parallelism the gradients (plus the KL statistic that drives the adaptive learning rate) are summed with
ONE NCCL all-reduce per minibatch step (dist.py).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.optim as optim

from .storage import RolloutStorage


class PPO:
    def __init__(self, actor_critic, num_learning_epochs=1, num_mini_batches=1, clip_param=0.2, gamma=0.998,
                 lam=0.95, value_loss_coef=1.0, entropy_coef=0.0, learning_rate=1e-3, max_grad_norm=1.0,
                 use_clipped_value_loss=True, schedule="fixed", desired_kl=0.01, device='cpu', weight_decay=0.0,
                 data_parallel=None):
        self.device = device
        self.desired_kl = desired_kl
        self.schedule = schedule
        self.learning_rate = learning_rate
        self.actor_critic = actor_critic
        self.actor_critic.to(self.device)
        self.storage = None
        self.optimizer = optim.Adam(self.actor_critic.parameters(), lr=learning_rate, weight_decay=weight_decay)
        self.transition = RolloutStorage.Transition()
        self.clip_param = clip_param
        self.num_learning_epochs = num_learning_epochs
        self.num_mini_batches = num_mini_batches
        self.value_loss_coef = value_loss_coef
        self.entropy_coef = entropy_coef
        self.gamma = gamma
        self.lam = lam
        self.max_grad_norm = max_grad_norm
        self.use_clipped_value_loss = use_clipped_value_loss
        self.dp = data_parallel            # dist.GradientAllReducer or None
        if self.dp is not None:
            self.dp.broadcast_parameters(self.actor_critic)

    def init_storage(self, num_envs, num_transitions_per_env, actor_obs_shape, critic_obs_shape, action_shape):
        self.storage = RolloutStorage(num_envs, num_transitions_per_env, actor_obs_shape, critic_obs_shape,
                                      action_shape, self.device)
        if self.dp is not None:
            self.storage.advantage_stats_allreduce = self.dp.allreduce_sum

    def test_mode(self):
        self.actor_critic.eval()

    def train_mode(self):
        self.actor_critic.train()

    def act(self, obs, critic_obs):
        ac = self.actor_critic
        self.transition.actions = ac.act(obs).detach()
        self.transition.values = ac.evaluate(critic_obs).detach()
        self.transition.actions_log_prob = ac.get_actions_log_prob(self.transition.actions).detach()
        self.transition.action_mean = ac.action_mean.detach()
        self.transition.action_sigma = ac.action_std.detach()
        self.transition.observations = obs
        self.transition.critic_observations = critic_obs
        return self.transition.actions

    def process_env_step(self, rewards, dones, infos):
        self.transition.rewards = rewards.clone()
        self.transition.dones = dones
        if 'time_outs' in infos:   # bootstrapping on time-outs (reference :108-109)
            self.transition.rewards += self.gamma * torch.squeeze(
                self.transition.values * infos['time_outs'].unsqueeze(1).to(self.device), 1)
        self.storage.add_transitions(self.transition)
        self.transition.clear()
        self.actor_critic.reset(dones)

    def compute_returns(self, last_critic_obs):
        last_values = self.actor_critic.evaluate(last_critic_obs).detach()
        self.storage.compute_returns(last_values, self.gamma, self.lam)

    # ------------------------------------------------------------------ update
    def _adapt_lr(self, kl_mean: float):
        if kl_mean > self.desired_kl * 2.0:
            self.learning_rate = max(1e-5, self.learning_rate / 1.5)
        elif kl_mean < self.desired_kl / 2.0 and kl_mean > 0.0:
            self.learning_rate = min(1e-2, self.learning_rate * 1.5)
        for param_group in self.optimizer.param_groups:
            param_group['lr'] = self.learning_rate

    def losses(self, batch):
        """Forward of one minibatch: returns (loss, surrogate_loss, value_loss, kl_mean tensor or None)."""
        (obs_batch, critic_obs_batch, actions_batch, target_values_batch, advantages_batch, returns_batch,
         old_actions_log_prob_batch, old_mu_batch, old_sigma_batch, hid_states_batch, masks_batch) = batch
        ac = self.actor_critic
        ac.act(obs_batch, masks=masks_batch, hidden_states=hid_states_batch[0])
        actions_log_prob_batch = ac.get_actions_log_prob(actions_batch)
        value_batch = ac.evaluate(critic_obs_batch, masks=masks_batch, hidden_states=hid_states_batch[1])
        mu_batch, sigma_batch, entropy_batch = ac.action_mean, ac.action_std, ac.entropy
        kl_mean = None
        if self.desired_kl is not None and self.schedule == 'adaptive':
            with torch.no_grad():
                kl = torch.sum(torch.log(sigma_batch / old_sigma_batch + 1.e-5)
                               + (torch.square(old_sigma_batch) + torch.square(old_mu_batch - mu_batch))
                               / (2.0 * torch.square(sigma_batch)) - 0.5, axis=-1)
                kl_mean = torch.mean(kl)
        ratio = torch.exp(actions_log_prob_batch - torch.squeeze(old_actions_log_prob_batch))
        surrogate = -torch.squeeze(advantages_batch) * ratio
        surrogate_clipped = -torch.squeeze(advantages_batch) * torch.clamp(ratio, 1.0 - self.clip_param,
                                                                           1.0 + self.clip_param)
        surrogate_loss = torch.max(surrogate, surrogate_clipped).mean()
        if self.use_clipped_value_loss:
            value_clipped = target_values_batch + (value_batch - target_values_batch).clamp(-self.clip_param,
                                                                                            self.clip_param)
            value_losses = (value_batch - returns_batch).pow(2)
            value_losses_clipped = (value_clipped - returns_batch).pow(2)
            value_loss = torch.max(value_losses, value_losses_clipped).mean()
        else:
            value_loss = (returns_batch - value_batch).pow(2).mean()
        loss = surrogate_loss + self.value_loss_coef * value_loss - self.entropy_coef * entropy_batch.mean()
        return loss, surrogate_loss, value_loss, kl_mean

    def update(self):
        generator = self.storage.mini_batch_generator(self.num_mini_batches, self.num_learning_epochs)
        vl_sum = torch.zeros((), device=self.device)
        sl_sum = torch.zeros((), device=self.device)
        for batch in generator:
            loss, surrogate_loss, value_loss, kl_mean = self.losses(batch)
            if kl_mean is not None:
                if self.dp is not None:
                    kl_mean = self.dp.allreduce_sum(kl_mean.reshape(1).double())[0] / self.dp.world_size
                self._adapt_lr(float(kl_mean))          # host sync: the reference's `if kl_mean > ...` (:145)
            self.optimizer.zero_grad()
            loss.backward()
            if self.dp is not None:
                self.dp.allreduce_gradients(self.actor_critic)
            nn.utils.clip_grad_norm_(self.actor_critic.parameters(), self.max_grad_norm)
            self.optimizer.step()
            vl_sum += value_loss.detach()
            sl_sum += surrogate_loss.detach()
        num_updates = self.num_learning_epochs * self.num_mini_batches
        self.storage.clear()
        return vl_sum.item() / num_updates, sl_sum.item() / num_updates
