"""PPO — same surface as rsl_rl.algorithms.PPO (reference: AMP/AMP_for_hardware-main/rsl_rl/rsl_rl/algorithms/
ppo.py:38-187; CASSIE identical but default lr, HUMANOID adds Adam weight_decay=1e-5 at :70-72).

act / process_env_step / compute_returns / update keep the reference's semantics.  Underneath, `update()` runs a
fused minibatch step with no host round trip (CUDA-graph replayed):

    actor forward (sm_100a kernels, traces kept)          snn_fwd_train
    critic forward (dense ELU MLP on the tensor cores,    snn_mlp_fwd     (a parallel branch of the CUDA graph;
    second stream)                                                         torch only if the critic is not Linear/ELU)
    PPO head: log-prob, entropy, KL, surrogate, value     snn_ppo_head    (forward + closed-form gradients)
    loss, and d loss / d (mu, value, log_std)
    actor BPTT backward straight into the flat bucket     snn_bwd
    critic backward                                       snn_mlp_bwd     (second stream)
    [data parallel: ONE all-reduce of the flat bucket, KL statistic riding in its tail]
    adaptive-KL learning rate on the device               snn_lr_adapt
    clip_grad_norm_ + Adam over the flat bucket           snn_clip_adam
    bf16 plane re-pack of the updated weights             snn_pack_weights

The generic path (`fused=False`, or a policy whose actor input needs a gradient such as the RMA latent) keeps
the reference's op-by-op structure through torch autograd, still on the same kernels.

CUDA graphs: a capture bakes in scalar kernel arguments (clip_param, value_loss_coef, entropy_coef, desired_kl, ...)
and buffer addresses.  The reference reads those attributes on every step, so `update()` compares a signature of all
of them (and of the storage's minibatch buffers) with the one the graphs were captured under and re-captures when
anything changed.  Every graph owns the workspaces its kernels write (never the engines' internal, re-allocatable
buffers).  A failed capture RAISES: a silent fall-back to eager launches would be a silent 3x slowdown
(SNN_GRAPH_FALLBACK=1 restores the print-and-continue behaviour for debugging).
"""
from __future__ import annotations

import os

import ctypes as C

import torch
import torch.nn as nn

from . import _lib
from .engine import CriticEngine
from .optim import FlatAdam
from .storage import RolloutStorage, RolloutStorageRMA


def _p(t):
    return C.c_void_p(t.data_ptr())


def _capture_failed(what, e):
    """A CUDA-graph capture failed: raise (default) — or, with SNN_GRAPH_FALLBACK=1, report and let the caller go on
    with eager launches."""
    if os.environ.get("SNN_GRAPH_FALLBACK", "0") == "1":
        print(f"[PPO] CUDA graph capture of {what} failed ({e}); SNN_GRAPH_FALLBACK=1: using eager launches")
        return
    raise RuntimeError(f"CUDA graph capture of {what} failed: {e}  (set SNN_GRAPH_FALLBACK=1 to continue with eager "
                       "launches, or use_cuda_graph=False)") from e


class _LazyNormal:
    """What `actor_critic.distribution` holds after a graph-replayed rollout step: Normal(mu, sigma) of that step,
    built on first use (constructing torch.distributions.Normal costs more host time than the replay itself)."""

    def __init__(self, mu, sigma):
        self.mean, self.stddev = mu, sigma
        self._d = None

    def __getattr__(self, name):           # log_prob, entropy, sample, ...
        if name.startswith("__") or name in ("_d", "mean", "stddev"):    # copy / pickle probe a bare instance
            raise AttributeError(name)
        if self._d is None:
            self._d = torch.distributions.Normal(self.mean, self.stddev, validate_args=False)
        return getattr(self._d, name)


class PPO:
    def __init__(self, actor_critic, num_learning_epochs=1, num_mini_batches=1, clip_param=0.2, gamma=0.998,
                 lam=0.95, value_loss_coef=1.0, entropy_coef=0.0, learning_rate=1e-3, max_grad_norm=1.0,
                 use_clipped_value_loss=True, schedule="fixed", desired_kl=0.01, device='cpu', weight_decay=0.0,
                 data_parallel=None, fused=True, use_cuda_graph=True, fused_critic=True):
        self.device = device
        self.desired_kl = desired_kl
        self.schedule = schedule
        self._lr = learning_rate
        self.actor_critic = actor_critic
        self.actor_critic.to(self.device)
        self.storage = None
        self.dp = data_parallel            # dist.GradientAllReducer or None
        if self.dp is not None:
            self.dp.broadcast_parameters(self.actor_critic)
        # `std` never receives a gradient in SNN mode (the actor owns log_std): torch's Adam skips it, so do we
        dead = {id(getattr(self.actor_critic, "std", None))}
        self._trainable = [p for p in self.actor_critic.parameters() if id(p) not in dead]
        self.optimizer = FlatAdam(self._trainable, lr=learning_rate, weight_decay=weight_decay, tail=1)
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
        self.fused = fused
        self.fused_critic = fused_critic
        self.act_cuda_graph = use_cuda_graph   # rollout step (act) replayed as one CUDA graph
        self._act_state = None
        # rollout step: the critic branch on the tensor-core MLP kernels (snn_mlp_fwd) instead of torch / cuBLAS SIMT
        # sgemm (SNN_ACT_CRITIC=torch: the torch branch)
        self.act_fused_critic = fused_critic and os.environ.get("SNN_ACT_CRITIC", "tc") != "torch"
        self._params_epoch = 0                 # bumped by every update(): the act graph's critic planes follow it
        # critic kernels on a second stream (parallel CUDA-graph branch); SNN_OVERLAP_CRITIC=0 serialises them
        self.overlap_critic = os.environ.get("SNN_OVERLAP_CRITIC", "1") != "0"
        # fused step: decoder + policy terms of the PPO head on the actor's chain, value loss on the critic's branch,
        # finalize on a third one (SNN_SPLIT_HEAD=0: the one-kernel head snn_ppo_head after a join of the two chains)
        self.split_head = os.environ.get("SNN_SPLIT_HEAD", "1") != "0"
        self.update_graph = os.environ.get("SNN_UPDATE_GRAPH", "1") != "0"      # one graph per update() (0: one per step)
        # capture the actor chain on a high-priority stream: where both chains have thread blocks pending, the
        # actor's (the critical path) are placed first (measured: 14.96 -> 14.61 ms per update; SNN_ACTOR_PRIO=0: off)
        self.actor_high_priority = os.environ.get("SNN_ACTOR_PRIO", "1") != "0"
        self._hp_stream = None
        self.use_cuda_graph = use_cuda_graph
        self._graphs = {}
        self._graph_launches = {}
        self._graph_sig = None                 # what the captured graphs baked in (_graph_signature)
        # data parallel: "single" = the NCCL all-reduce is captured inside the step graph, "segments" = two graph
        # segments around an eager all-reduce (see _run_fast_dp)
        self.dp_graph_mode = os.environ.get("SNN_DP_GRAPH", "single")
        # data parallel on CUDA: the sum of the gradient bucket over the ranks inside the optimizer kernel, over NVLink
        # peer memory (all ranks set it up together; False = NCCL all-reduce between backward and optimizer)
        self.dp_p2p = bool(self.dp is not None and self.optimizer.flat_params.is_cuda and hasattr(self.dp, "enable_p2p")
                           and self.dp.enable_p2p(self.optimizer.flat_grads.numel()))
        self._p2p_obj = self.dp.p2p if self.dp_p2p else None      # this instance's buffers (its graphs bake the pointers in)
        self.dp_graph_note = ""
        self.replayed_kernel_launches = 0      # kernels of this library launched through CUDA-graph replays
        self._fast_state = None
        dev = self.optimizer.flat_params.device
        self._stats = torch.zeros(8, dtype=torch.float32, device=dev)

    # learning_rate is a plain attribute in the reference; here the truth lives on the device during update()
    @property
    def learning_rate(self):
        return self._lr

    @learning_rate.setter
    def learning_rate(self, v):
        self._lr = float(v)
        if hasattr(self, "optimizer"):
            self.optimizer.set_lr(self._lr)

    def init_storage(self, num_envs, num_transitions_per_env, actor_obs_shape, critic_obs_shape, action_shape):
        self.storage = RolloutStorage(num_envs, num_transitions_per_env, actor_obs_shape, critic_obs_shape,
                                      action_shape, self.device)
        if self.dp is not None:
            self.storage.advantage_stats_allreduce = self.dp.allreduce_sum
        self._graphs, self._graph_launches, self._graph_sig = {}, {}, None      # graphs read the old storage's buffers

    def test_mode(self):
        self.actor_critic.eval()

    def train_mode(self):
        self.actor_critic.train()

    def _act_body(self, obs, critic_obs):
        ac = self.actor_critic
        actions = ac.act(obs).detach()
        values = ac.evaluate(critic_obs).detach()
        logp = ac.get_actions_log_prob(actions).detach()
        return actions, values, logp, ac.action_mean.detach(), ac.action_std.detach()

    def _act_fused(self, st):
        """Graph-safe rollout step on static buffers: the critic on a parallel branch (second stream), the actor
        forward writing mu in place, then ONE kernel for sigma / sampling / log-prob (snn_sample_logprob; the
        reference's Normal.sample + log_prob + stddev are ~16 elementwise launches).  torch.normal(mean, std) would
        check `std >= 0` on the host (a sync, not capturable); the noise comes from randn_like instead — same
        distribution, parity is defined on mu / sigma / values / log-probs, not on the random draw."""
        ac = self.actor_critic
        actor = ac.actor
        dev = st["obs"].device
        o = st["views"]
        cur = torch.cuda.current_stream(dev)
        side = st["side"]
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            if st.get("ce") is not None:
                st["ce"].forward(st["cobs"], out_value=o["values"].view(-1), out_trace=st["ctrace"])
            else:
                o["values"].copy_(ac.evaluate(st["cobs"]))
        layers = actor.snn.linear_layers()
        # the parameters themselves (not .data: that hides torch's version counters from the engine's "re-pack only
        # if a weight changed" check; _act_graphed runs the same check on the host before every replay)
        params = (actor.encoder.mean, actor.encoder.std, [l.weight for l in layers], [l.bias for l in layers],
                  actor.decoder.decoder.weight, actor.decoder.decoder.bias)
        B, A = o["mu"].shape
        noise = torch.randn_like(o["mu"])
        if actor.engine.act_preferred(B) and A <= 64:
            # ONE launch: encoder, LIF layers, decoder, sigma / sampling / log-prob (snn_fwd_act)
            actor.engine.act(st["obs"], *params, noise, actor.log_std.data, o["mu"], o["actions"], o["logp"], o["sigma"])
        else:
            actor.engine.forward(st["obs"], *params, train=False, out_mu=o["mu"], ws=st["ws_fwd"])
            with torch.cuda.device(dev):
                _lib.check(_lib.lib().snn_sample_logprob(_p(o["mu"]), _p(actor.log_std.data), _p(noise), B, A,
                                                         _p(o["actions"]), _p(o["logp"]), _p(o["sigma"]),
                                                         C.c_void_p(cur.cuda_stream)), "snn_sample_logprob")
        cur.wait_stream(side)

    @staticmethod
    def _act_views(packed, B, A):
        off = [0, B * A, B * A + B, B * A + 2 * B, 2 * B * A + 2 * B]
        return {"actions": packed[off[0]:off[1]].view(B, A), "values": packed[off[1]:off[2]].view(B, 1),
                "logp": packed[off[2]:off[3]], "mu": packed[off[3]:off[4]].view(B, A),
                "sigma": packed[off[4]:off[4] + B * A].view(B, A)}

    def _act_graphed(self, obs, critic_obs):
        """Rollout step as ONE CUDA-graph replay (the reference launches ~250 kernels here): inputs are copied into
        static buffers; actions / values / log-probs / mu / sigma land in one packed static buffer that is cloned once
        per call (fresh tensors for the caller, as in the reference)."""
        key = (tuple(obs.shape), tuple(critic_obs.shape), obs.data_ptr() == critic_obs.data_ptr(), obs.dtype,
               critic_obs.dtype)
        st = self._act_state
        if st is None or st["key"] != key:
            B, A = obs.shape[0], self.actor_critic.actor.act_dim
            dev = obs.device
            # static buffers are plain (non-inference) contiguous tensors whatever mode / strides the first call came
            # with: a later call under torch.no_grad() or with a strided view must still be able to copy_ into them
            with torch.inference_mode(False):
                st = {"key": key, "graph": None, "calls": 0,
                      "obs": torch.empty(obs.shape, dtype=obs.dtype, device=dev),
                      "packed": torch.empty(3 * B * A + 2 * B, dtype=torch.float32, device=dev),
                      "ws_fwd": torch.empty(max(self.actor_critic.actor.engine.workspace_bytes(B, False), 1024),
                                            dtype=torch.uint8, device=dev),
                      "side": torch.cuda.Stream(device=dev)}
                st["cobs"] = st["obs"] if key[2] else torch.empty(critic_obs.shape, dtype=critic_obs.dtype, device=dev)
                st["views"] = self._act_views(st["packed"], B, A)
                st["ce"] = None
                if self.act_fused_critic and isinstance(self.actor_critic.critic, nn.Sequential) \
                        and st["cobs"].dtype == torch.float32:
                    ce = CriticEngine.from_sequential(self.actor_critic.critic)
                    if ce is not None:
                        st["ce"], st["ce_ver"] = ce, None
                        st["ctrace"] = torch.empty(ce.trace_bytes(B), dtype=torch.uint8, device=dev)
            self._act_state = st
        st["obs"].copy_(obs)
        if not key[2]:
            st["cobs"].copy_(critic_obs)
        st["calls"] += 1
        ce = st["ce"]
        if ce is not None:
            # operand planes of the critic: re-derived when an update() ran or a parameter was written through torch
            ver = (self._params_epoch, sum(l.weight._version + l.bias._version for l in ce.linears))
            if st["ce_ver"] != ver:
                ce.repack()
                st["ce_ver"] = ver
        if st["graph"] is None:
            if st["calls"] < 2:                       # first call eager (lazy init, weight packing)
                return self._act_body(st["obs"], st["cobs"])
            side = torch.cuda.Stream()                # warm-up on a side stream, as torch.cuda.graph requires
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    self._act_fused(st)               # kernel / argument errors surface here, un-caught
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            try:
                with torch.cuda.graph(g):
                    self._act_fused(st)
            except Exception as e:  # noqa: BLE001   (capture-time failure only)
                _capture_failed("PPO.act", e)
                self.act_cuda_graph = False
                return self._act_body(st["obs"], st["cobs"])
            st["graph"] = g
        actor = self.actor_critic.actor
        layers = st.get("layers")
        if layers is None:
            layers = st["layers"] = actor.snn.linear_layers()
        actor.engine.ensure_packed(actor.encoder.mean, actor.encoder.std, [l.weight for l in layers],
                                   [l.bias for l in layers], actor.decoder.decoder.weight, actor.decoder.decoder.bias)
        st["graph"].replay()
        B, A = st["views"]["mu"].shape
        v = self._act_views(st["packed"].clone(), B, A)
        self.actor_critic.distribution = _LazyNormal(v["mu"], v["sigma"])     # reference: act() leaves it set
        return v["actions"], v["values"], v["logp"], v["mu"], v["sigma"]

    def act(self, obs, critic_obs):
        use_graph = (self.act_cuda_graph and obs.is_cuda and not torch.is_grad_enabled()
                     and type(self.actor_critic).__name__ != "ActorCriticRMA" and self.actor_critic.actor.act_dim <= 64)
        out = self._act_graphed(obs, critic_obs) if use_graph else self._act_body(obs, critic_obs)
        (self.transition.actions, self.transition.values, self.transition.actions_log_prob,
         self.transition.action_mean, self.transition.action_sigma) = out
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

    # ------------------------------------------------------------------ generic (autograd) minibatch step
    def _loss_terms(self, actions_batch, target_values_batch, advantages_batch, returns_batch,
                    old_actions_log_prob_batch, old_mu_batch, old_sigma_batch, value_batch):
        """The reference's loss block (ppo.py:135-171) on the distribution `actor_critic` currently holds:
        (loss, surrogate_loss, value_loss, kl_mean or None)."""
        ac = self.actor_critic
        actions_log_prob_batch = ac.get_actions_log_prob(actions_batch)
        mu_batch, sigma_batch, entropy_batch = ac.action_mean, ac.action_std, ac.entropy
        kl_mean = None
        if self._adaptive():
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

    def losses(self, batch):
        """Forward of one minibatch through autograd: (loss, surrogate_loss, value_loss, kl_mean or None)."""
        (obs_batch, critic_obs_batch, actions_batch, target_values_batch, advantages_batch, returns_batch,
         old_actions_log_prob_batch, old_mu_batch, old_sigma_batch, hid_states_batch, masks_batch) = batch
        ac = self.actor_critic
        ac.act(obs_batch, masks=masks_batch, hidden_states=hid_states_batch[0])
        value_batch = ac.evaluate(critic_obs_batch, masks=masks_batch, hidden_states=hid_states_batch[1])
        return self._loss_terms(actions_batch, target_values_batch, advantages_batch, returns_batch,
                                old_actions_log_prob_batch, old_mu_batch, old_sigma_batch, value_batch)

    def _adaptive(self):
        return self.desired_kl is not None and self.schedule == 'adaptive'

    def _p2p(self):
        """The peer-memory gradient sum of a data-parallel CUDA job (dist.P2PGradientSum), or None: single process, CPU /
        gloo, SNN_DP_P2P=0, or the IPC mapping failed (then every rank falls back to the NCCL all-reduce alike)."""
        return self._p2p_obj if self.dp_p2p else None

    def _finish_step(self):
        """all-reduce (data parallel) -> adaptive lr -> clip + Adam -> re-pack: shared by both paths.  With peer memory
        the sum over the ranks is part of the optimizer kernel (snn_clip_adam_kl_p2p): no collective call at all."""
        if self.dp is not None and self._p2p() is None:
            self.dp.allreduce_sum(self.optimizer.flat_grads)   # ONE collective: gradients + KL statistic in the tail
        self._post_allreduce()

    def _post_allreduce(self):
        opt = self.optimizer
        world = 1 if self.dp is None else self.dp.world_size
        dev = opt.flat_params.device
        p2p = self._p2p()
        if p2p is not None:
            adaptive = self._adaptive()
            opt.step_p2p(p2p, max_grad_norm=self.max_grad_norm if self.max_grad_norm is not None else 0.0,
                         grad_scale=1.0 / world, adaptive=adaptive, kl_scale=1.0 / world,
                         desired_kl=float(self.desired_kl) if adaptive else 0.0)
            self._repack_actor()
            return
        # adaptive lr (ppo.py:145-148) + clip + Adam: one kernel (the KL statistic sits in the tail of the gradient bucket)
        adaptive = self._adaptive()
        opt.step(max_grad_norm=self.max_grad_norm if self.max_grad_norm is not None else 0.0, grad_scale=1.0 / world,
                 kl=opt.flat_grads[opt.n:] if adaptive else None, kl_scale=1.0 / world,
                 desired_kl=float(self.desired_kl) if adaptive else 0.0)
        self._repack_actor()
        # the critic's operand planes are re-derived at the head of the next step's critic branch (second stream):
        # off the critical path, which continues with the actor's encoder (_step_fast)

    def _repack_actor(self):
        """bf16 operand planes <- fp32 parameters, unconditionally (the fused optimizer writes parameters behind
        torch's version counters).  The Parameters themselves are passed, so the engine's "changed since the last pack"
        key stays the one `ensure_packed` computes before a rollout-graph replay: no second re-pack there."""
        actor = self.actor_critic.actor
        layers = actor.snn.linear_layers()
        actor.engine.repack(actor.encoder.mean, actor.encoder.std, [l.weight for l in layers],
                            [l.bias for l in layers], actor.decoder.decoder.weight, actor.decoder.decoder.bias)

    def _step_generic(self, batch):
        loss, surrogate_loss, value_loss, kl_mean = self.losses(batch)
        opt = self.optimizer
        opt.zero_grad()
        loss.backward()
        if kl_mean is not None:
            opt.flat_grads[opt.n:].copy_(kl_mean.reshape(1))
        self._finish_step()
        self._stats[4] += surrogate_loss.detach()
        self._stats[5] += value_loss.detach()

    # ------------------------------------------------------------------ fused minibatch step
    def _fast_supported(self):
        ac = self.actor_critic
        return (self.fused and type(ac).__name__ != "ActorCriticRMA" and not ac.is_recurrent
                and self.optimizer.flat_params.is_cuda and ac.actor.act_dim <= 64)

    def _graph_signature(self, f, mb):
        """Everything a captured minibatch-step graph bakes in: the scalar kernel arguments the reference reads from
        `self` on every step, and the addresses of the minibatch buffers.  `update()` re-captures when it changes (a
        runner that anneals entropy_coef / clip_param, switches `schedule`, or calls init_storage() again)."""
        g = self.optimizer.param_groups[0]
        return (mb, float(self.clip_param), float(self.value_loss_coef), float(self.entropy_coef),
                bool(self.use_clipped_value_loss), self.desired_kl, self.schedule, self.max_grad_norm,
                tuple(g["betas"]), float(g["eps"]), float(g["weight_decay"]), self.overlap_critic, self.split_head,
                self.update_graph, self.num_learning_epochs, self.num_mini_batches,
                1 if self.dp is None else self.dp.world_size, self.dp_p2p,
                tuple(sorted((k, v.data_ptr()) for k, v in f.items())))

    def _fast_setup(self, mb):
        ac = self.actor_critic
        dev = self.optimizer.flat_params.device
        A = ac.actor.act_dim
        eng = ac.actor.engine
        st = {
            "mb": mb,
            "mu": torch.empty(mb, A, dtype=torch.float32, device=dev),
            "g_mu": torch.empty(mb, A, dtype=torch.float32, device=dev),
            "g_value": torch.empty(mb, dtype=torch.float32, device=dev),
            "trace": torch.empty(eng.trace_bytes(mb), dtype=torch.uint8, device=dev),
            # owned by this state (and so by the graphs captured over it), not by the engine: an eager call with a
            # larger batch re-allocates the engine's internal workspace, which a replayed graph would still write to
            "ws_bwd": torch.empty(eng.workspace_bytes(mb, True), dtype=torch.uint8, device=dev),
            "scratch": torch.empty(((mb + 255) // 256) * (4 + A) * 4 + 1024, dtype=torch.uint8, device=dev),
        }
        layers = ac.actor.snn.linear_layers()
        st["actor_tensors"] = (ac.actor.encoder.mean, ac.actor.encoder.std, [l.weight for l in layers],
                               [l.bias for l in layers], ac.actor.decoder.decoder.weight, ac.actor.decoder.decoder.bias)
        st["grads_out"] = {
            "enc_mean": ac.actor.encoder.mean.grad, "enc_std": ac.actor.encoder.std.grad,
            "weights": [l.weight.grad for l in layers], "biases": [l.bias.grad for l in layers],
            "dec_w": ac.actor.decoder.decoder.weight.grad, "dec_b": ac.actor.decoder.decoder.bias.grad, "obs": None,
        }
        st["head_part"] = torch.empty(max(eng.ppo_head_part_bytes(mb), 1024), dtype=torch.uint8, device=dev)
        st["value_part"] = torch.empty(((mb + 255) // 256) * 4 + 1024, dtype=torch.uint8, device=dev)
        st["critic"] = None
        if self.fused_critic:
            ce = CriticEngine.from_sequential(ac.critic) if isinstance(ac.critic, nn.Sequential) else None
            if ce is not None:
                ce.repack()
                st["critic"] = ce
                st["value"] = torch.empty(mb, dtype=torch.float32, device=dev)
                st["critic_trace"] = torch.empty(ce.trace_bytes(mb), dtype=torch.uint8, device=dev)
                st["critic_ws"] = torch.empty(ce.workspace_bytes(mb), dtype=torch.uint8, device=dev)
                st["critic_grads"] = ([l.weight.grad for l in ce.linears], [l.bias.grad for l in ce.linears])
                st["side_stream"] = torch.cuda.Stream(device=dev) if self.overlap_critic else None
                st["fin_stream"] = torch.cuda.Stream(device=dev) if self.overlap_critic else None
        self._fast_state = st
        self._graphs = {}

    def _step_fast(self, f, i, finish=True):
        """One fused minibatch step on rows [i*mb, (i+1)*mb) of the permuted fields `f`.  No host sync.
        finish=False stops before the all-reduce / optimizer (data-parallel CUDA-graph segments)."""
        st = self._fast_state
        mb = st["mb"]
        sl = slice(i * mb, (i + 1) * mb)
        ac, opt = self.actor_critic, self.optimizer
        dev = opt.flat_params.device
        em, es, ws, bs, dw, db = st["actor_tensors"]
        eng = ac.actor.engine
        ce = st["critic"]
        # The gradient bucket is cleared per step only on the paths that ACCUMULATE into it (torch autograd for a critic
        # that is not on the fused kernels).  Every fused kernel overwrites its slice (actor: snn_bwd*, critic: snn_mlp_bwd,
        # log_std and the KL tail: the PPO head's finalize), and update() clears the bucket once per call for parameters no
        # kernel writes (the reference's unused `std`).  The per-step fill sat on the critical path between the re-pack
        # and the next step's encoder.
        if ce is None:
            opt.zero_grad()
        side = st.get("side_stream") if ce is not None else None
        main = torch.cuda.current_stream(dev)
        if ce is not None:
            # the critic chain is independent of the actor chain until the PPO head: run it on a second stream so
            # that its kernels fill the launch / drain gaps of the actor's persistent kernels (fork/join is captured
            # into the CUDA graph as parallel branches)
            if side is not None:
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    ce.repack()
                    value, ctrace = ce.forward(f["critic_obs"][sl], out_value=st["value"], out_trace=st["critic_trace"])
            else:
                ce.repack()
                value, ctrace = ce.forward(f["critic_obs"][sl], out_value=st["value"], out_trace=st["critic_trace"])
        split = ce is not None and self.split_head
        params = (em.data, es.data, [w.data for w in ws], [b.data for b in bs], dw.data, db.data)
        mu, trace = eng.forward(f["obs"][sl], *params, train=True, out_mu=False if split else st["mu"],
                                out_trace=st["trace"], assume_packed=True)
        L = _lib.lib()
        world = 1 if self.dp is None else self.dp.world_size
        if split:
            # The policy terms of the PPO head need nothing from the critic: the actor's chain runs decoder + head as one
            # kernel, then the output layer's scan (snn_bwd_ppo_top) and goes straight on with its dX chain; the value loss
            # stays on the critic's branch, and the finalize (loss statistics, d log_std, KL into the tail of the gradient
            # bucket for the all-reduce / adaptive lr) joins the two partial sets on a third branch, off both chains.
            eng.backward_ppo_top(f["obs"][sl], trace, params, ac.actor.log_std.data, f["actions"][sl], f["logp"][sl],
                                 f["adv"][sl], f["mu"][sl], f["sigma"][sl], float(self.clip_param), st["head_part"],
                                 st["ws_bwd"], out_mu=st["mu"], out_g_mu=st["g_mu"])
            def finalize(stream_):
                _lib.check(L.snn_ppo_finalize(C.byref(eng.desc), mb, _p(st["head_part"]), _p(st["value_part"]),
                                              _p(ac.actor.log_std.data), float(self.entropy_coef), _p(ac.actor.log_std.grad),
                                              _p(self._stats), _p(opt.flat_grads[opt.n:]), C.c_void_p(stream_.cuda_stream)),
                           "snn_ppo_finalize")

            branch = side if side is not None else main
            head_done = vl_done = None
            if side is not None:
                head_done = torch.cuda.Event()
                head_done.record(main)
            with torch.cuda.stream(branch), torch.cuda.device(dev):
                _lib.check(L.snn_ppo_value_loss(_p(value.detach()), _p(f["returns"][sl]), _p(f["values"][sl]), mb,
                                                float(self.clip_param), float(self.value_loss_coef),
                                                1 if self.use_clipped_value_loss else 0, _p(st["g_value"]),
                                                _p(st["value_part"]), st["value_part"].numel(),
                                                C.c_void_p(branch.cuda_stream)), "snn_ppo_value_loss")
                if side is not None:
                    vl_done = torch.cuda.Event()
                    vl_done.record(side)
                else:
                    finalize(main)
                ce.backward(st["g_value"], ctrace, mb, out=st["critic_grads"], ws=st["critic_ws"])
            if side is not None:
                # third branch: the finalize needs the head's partial sums and the value loss only — neither the actor's
                # dX chain nor the critic's backward wait for it, and it is done long before the optimizer joins them
                fin = st["fin_stream"]
                fin.wait_event(head_done)
                fin.wait_event(vl_done)
                with torch.cuda.stream(fin), torch.cuda.device(dev):
                    finalize(fin)
            eng.backward_rest(f["obs"][sl], trace, params, st["grads_out"], st["ws_bwd"])
            if side is not None:
                main.wait_stream(side)
                main.wait_stream(fin)
            if finish:
                self._finish_step()
            return
        if ce is None:
            with torch.enable_grad():
                value = ac.evaluate(f["critic_obs"][sl])
        elif side is not None:
            main.wait_stream(side)
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        with torch.cuda.device(dev):
            _lib.check(L.snn_ppo_head(
                _p(mu), _p(ac.actor.log_std.data), _p(value.detach()), _p(f["actions"][sl]), _p(f["logp"][sl]),
                _p(f["adv"][sl]), _p(f["returns"][sl]), _p(f["values"][sl]), _p(f["mu"][sl]), _p(f["sigma"][sl]),
                mb, ac.actor.act_dim, float(self.clip_param), float(self.value_loss_coef), float(self.entropy_coef),
                1 if self.use_clipped_value_loss else 0, _p(st["g_mu"]), _p(st["g_value"]), _p(ac.actor.log_std.grad),
                _p(self._stats), _p(st["scratch"]), st["scratch"].numel(), stream), "snn_ppo_head")
        # KL statistic rides in the tail of the gradient bucket (one all-reduce under data parallelism)
        opt.flat_grads[opt.n:].copy_(self._stats[2:3])
        if ce is not None and side is not None:
            side.wait_stream(main)
            with torch.cuda.stream(side):
                ce.backward(st["g_value"], ctrace, mb, out=st["critic_grads"], ws=st["critic_ws"])
        eng.backward(f["obs"][sl], mu, st["g_mu"], trace, em.data, es.data, [w.data for w in ws],
                     [b.data for b in bs], dw.data, db.data, need_dobs=False, out=st["grads_out"], assume_packed=True,
                     ws=st["ws_bwd"])
        if ce is not None:
            if side is not None:
                main.wait_stream(side)
            else:
                ce.backward(st["g_value"], ctrace, mb, out=st["critic_grads"], ws=st["critic_ws"])
        else:
            value.backward(st["g_value"].view_as(value))
        if finish:
            self._finish_step()

    def _capture(self, fn, snap, what):
        """Warm `fn` up on a side stream, restore the training state, capture it into a CUDA graph.
        Returns (graph, kernel launches of this library inside it) or None when the capture failed and
        SNN_GRAPH_FALLBACK=1 allows going on without it."""
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            fn()                                   # kernel / argument errors surface here, un-caught
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self._restore(snap)
        g = torch.cuda.CUDAGraph()
        n0 = _lib.lib().snn_launch_count()
        try:
            with torch.cuda.graph(g, **self._graph_kw()):
                fn()
        except Exception as e:  # noqa: BLE001   (capture-time failure only)
            self._restore(snap)
            _capture_failed(what, e)
            return None
        self._restore(snap)
        return g, int(_lib.lib().snn_launch_count() - n0)

    def _graph_kw(self):
        if not self.actor_high_priority:
            return {}
        if self._hp_stream is None:
            self._hp_stream = torch.cuda.Stream(device=self.optimizer.flat_params.device, priority=-1)
        return {"stream": self._hp_stream}

    def _run_fast_dp(self, f, i):
        """Data parallel.  dp_graph_mode "single": the whole step INCLUDING the NCCL all-reduce of the flat bucket is
        one CUDA graph (NCCL's kernel is captured like any other; no graph boundary, no eager NCCL launch on the
        critical path).  "segments": graph A (everything up to the local gradients) -> eager NCCL all-reduce ->
        graph B (adaptive lr, clip + Adam, re-pack) — the round-1 scheme, kept for NCCL builds that cannot be
        captured (selected with SNN_DP_GRAPH=segments, or automatically when the single-graph capture fails; the
        mode that ran is reported by bench.py as `dp_graph`)."""
        if self.dp_graph_mode == "single":
            if ("full", i) not in self._graphs:
                try:
                    cap = self._capture(lambda: self._step_fast(f, i), self._snapshot(), "the data-parallel PPO step")
                except RuntimeError as e:
                    cap = None
                    self.dp_graph_mode, self.dp_graph_note = "segments", f"single-graph capture failed: {e}"
                if cap is not None:
                    self._graphs[("full", i)], self._graph_launches[("full", i)] = cap
            if ("full", i) in self._graphs:
                self._graphs[("full", i)].replay()
                self.replayed_kernel_launches += self._graph_launches[("full", i)]
                return
            self.dp_graph_mode = "segments"
        if ("pre", i) not in self._graphs:
            cap = self._capture(lambda: self._step_fast(f, i, finish=False), self._snapshot(), "the PPO step (pre)")
            if cap is None:
                self.use_cuda_graph = False
                return self._step_fast(f, i)
            self._graphs[("pre", i)], self._graph_launches[("pre", i)] = cap
        if "post" not in self._graphs:
            gsnap = self.optimizer.flat_grads.clone()
            cap = self._capture(self._post_allreduce, self._snapshot(), "the PPO step (post)")
            self.optimizer.flat_grads.copy_(gsnap)
            if cap is None:
                self.use_cuda_graph = False
                return self._step_fast(f, i)
            self._graphs["post"], self._graph_launches["post"] = cap
        self._graphs[("pre", i)].replay()
        if self._p2p() is None:        # peer-memory sum: part of the optimizer kernel in the "post" graph
            self.dp.allreduce_sum(self.optimizer.flat_grads)
        self._graphs["post"].replay()
        self.replayed_kernel_launches += self._graph_launches[("pre", i)] + self._graph_launches["post"]

    def _run_fast(self, f, i):
        if not self.use_cuda_graph:
            return self._step_fast(f, i)
        if self.dp is not None:
            return self._run_fast_dp(f, i)
        if i not in self._graphs:
            cap = self._capture(lambda: self._step_fast(f, i), self._snapshot(), "the PPO step")
            if cap is None:
                self.use_cuda_graph = False
                return self._step_fast(f, i)
            self._graphs[i], self._graph_launches[i] = cap
        self._graphs[i].replay()
        self.replayed_kernel_launches += self._graph_launches[i]

    def _run_update_graph(self, f):
        """Single process: ALL minibatch steps of the update (epochs x minibatches) as ONE CUDA graph — between two
        per-step graph replays the GPU idles for 5-8 us (the next graph's launch), 19 times per update at C2.  The steps
        read the permuted fields `f` in place, so the same graph serves every update until the signature changes.
        Returns False when this mode does not apply (eager mode, data parallel: per-step graphs around the all-reduce)."""
        if not self.use_cuda_graph or (self.dp is not None and self._p2p() is None) or not self.update_graph:
            return False      # NCCL inside the graph: one graph per step (_run_fast_dp), as validated in r2
        if "update" not in self._graphs:
            def all_steps():
                for _ in range(self.num_learning_epochs):
                    for i in range(self.num_mini_batches):
                        self._step_fast(f, i)
            cap = self._capture(all_steps, self._snapshot(), "the PPO update")
            if cap is None:
                self.use_cuda_graph = False
                return False
            self._graphs["update"], self._graph_launches["update"] = cap
        self._graphs["update"].replay()
        self.replayed_kernel_launches += self._graph_launches["update"]
        return True

    def _snapshot(self):
        opt = self.optimizer
        return (opt.flat_params.clone(), opt.exp_avg.clone(), opt.exp_avg_sq.clone(), opt.step_dev.clone(),
                opt.lr_dev.clone(), self._stats.clone())

    def _restore(self, snap):
        opt = self.optimizer
        opt.flat_params.copy_(snap[0]); opt.exp_avg.copy_(snap[1]); opt.exp_avg_sq.copy_(snap[2])
        opt.step_dev.copy_(snap[3]); opt.lr_dev.copy_(snap[4]); self._stats.copy_(snap[5])
        self.refresh_packed()

    def refresh_packed(self):
        """Re-derive the bf16 operand planes from the fp32 parameters (after anything that changed them behind the
        fused path's back: a restored snapshot, load_state_dict / load_checkpoint, a write through `param.data`)."""
        self._repack_actor()
        if self._fast_state is not None and self._fast_state.get("critic") is not None:
            self._fast_state["critic"].repack()

    # ------------------------------------------------------------------ update
    def update(self):
        self._stats.zero_()
        self.optimizer.set_lr(self._lr)
        num_updates = self.num_learning_epochs * self.num_mini_batches
        if self._fast_supported():
            f, mb = self.storage.permute_once(self.num_mini_batches)
            if self._fast_state is None or self._fast_state["mb"] != mb:
                self._fast_setup(mb)
            sig = self._graph_signature(f, mb)
            if sig != self._graph_sig:                 # a hyperparameter or a buffer address changed: re-capture
                self._graphs, self._graph_launches, self._graph_sig = {}, {}, sig
            self.refresh_packed()      # the fused steps assume current planes (parameters may have been loaded since)
            self.optimizer.zero_grad()     # once per update: see _step_fast
            if not self._run_update_graph(f):
                for _ in range(self.num_learning_epochs):
                    for i in range(self.num_mini_batches):
                        self._run_fast(f, i)
        else:
            for batch in self.storage.mini_batch_generator(self.num_mini_batches, self.num_learning_epochs):
                self._step_generic(batch)
        stats = self._stats.tolist()                       # the one host sync of the update
        self._lr = self.optimizer.sync_lr_from_device()
        self._params_epoch += 1
        self.storage.clear()
        return stats[5] / num_updates, stats[4] / num_updates


class PPORMA(PPO):
    """RMA fork's PPO (RMA/quadruped-robot-master/rl-robotics/rsl_rl/rsl_rl/algorithms/ppo.py:41-221): the policy acts
    on (obs, privileged_obs), and every minibatch step is followed by `num_adaptation_module_substeps` regression
    steps of the adaptation module onto the (detached) privileged-encoder latent with a second Adam (:200-213).

    Fused path (default): one CUDA graph per minibatch index holding
        env_factor_encoder forward (tensor-core MLP kernels, snn_mlp_fwd_ex) straight into the latent columns of the
        cat(obs, latent) buffer -> spiking actor forward + dense critic forward on it -> snn_ppo_head -> actor BPTT with
        d loss / d actor-input (snn_bwd) and critic backward with d loss / d critic-input (snn_mlp_bwd_ex) -> encoder
        backward on the SUM of the two latent gradients -> all-reduce / adaptive lr / clip + Adam -> the adaptation
        substeps (adaptation_module forward, encoder target forward, snn_mse_grad, backward, second Adam over the
        adaptation module's slice of the flat parameter buffer through snn_clip_adam).
    `fused=False` keeps the autograd path (torch MLPs) on the same spiking kernels."""

    def __init__(self, actor_critic, num_adaptation_module_substeps=1, **kwargs):
        super().__init__(actor_critic, **kwargs)
        self.num_adaptation_module_substeps = num_adaptation_module_substeps
        self._adaptation_lr = float(kwargs.get("learning_rate", 1e-3))
        self.adaptation_optimizer = torch.optim.Adam(self.actor_critic.adaptation_module.parameters(),
                                                     lr=self._adaptation_lr)
        self._ad = None                         # fused path: Adam state of the adaptation module's flat slice
        self._rma_fusable = None
        dev = self.optimizer.flat_params.device
        self._adapt_sum = torch.zeros(1, dtype=torch.float32, device=dev)

    # ------------------------------------------------------------------ fused minibatch step (RMA)
    def _fast_supported(self):
        ac = self.actor_critic
        if not (self.fused and self.fused_critic and not ac.is_recurrent and self.optimizer.flat_params.is_cuda
                and ac.actor.act_dim <= 64):
            return False
        if self._rma_fusable is None:
            from .engine import MlpEngine
            self._rma_fusable = all(isinstance(m, nn.Sequential) for m in (ac.env_factor_encoder, ac.adaptation_module,
                                                                           ac.critic)) \
                and MlpEngine.from_sequential(ac.env_factor_encoder) is not None \
                and MlpEngine.from_sequential(ac.adaptation_module) is not None \
                and CriticEngine.from_sequential(ac.critic) is not None
        return self._rma_fusable

    def _fast_setup(self, mb):
        from .engine import MlpEngine
        super()._fast_setup(mb)
        st, ac, opt = self._fast_state, self.actor_critic, self.optimizer
        dev = opt.flat_params.device
        ce = st["critic"]
        assert ce is not None
        enc = MlpEngine.from_sequential(ac.env_factor_encoder)
        ad = MlpEngine.from_sequential(ac.adaptation_module)
        D, Lz = int(ce.desc.in_dim), enc.out_dim
        f32 = dict(dtype=torch.float32, device=dev)
        u8 = dict(dtype=torch.uint8, device=dev)
        st.update({
            "O": D - Lz, "Lz": Lz,
            "x": torch.zeros(mb, D, **f32),                 # cat(obs, latent): the actor's and the critic's input
            "dx_actor": torch.zeros(mb, D, **f32), "dx_critic": torch.zeros(mb, D, **f32),
            "enc": enc, "enc_trace": torch.empty(enc.trace_bytes(mb), **u8), "enc_ws": torch.empty(enc.workspace_bytes(mb), **u8),
            "enc_grads": ([l.weight.grad for l in enc.linears], [l.bias.grad for l in enc.linears]),
            "ad": ad, "ad_trace": torch.empty(ad.trace_bytes(mb), **u8), "ad_ws": torch.empty(ad.workspace_bytes(mb), **u8),
            "ad_grads": ([l.weight.grad for l in ad.linears], [l.bias.grad for l in ad.linears]),
            "ad_pred": torch.empty(mb, Lz, **f32), "ad_tgt": torch.empty(mb, Lz, **f32), "ad_g": torch.empty(mb, Lz, **f32),
            "ad_scratch": torch.empty(4096, **u8),
        })
        st["grads_out"]["obs"] = st["dx_actor"]
        enc.repack(); ad.repack()
        if self._ad is None:
            # the adaptation module's parameters are one contiguous slice of the flat buffer (parameters() walks modules)
            params = list(ac.adaptation_module.parameters())
            off = (params[0].data_ptr() - opt.flat_params.data_ptr()) // 4
            k, cur = 0, params[0].data_ptr()
            for p in params:
                if p.data_ptr() != cur:
                    raise RuntimeError("adaptation_module parameters are not contiguous in the flat parameter buffer")
                cur += p.numel() * 4
                k += p.numel()
            self._ad = {"off": off, "k": k, "m": torch.zeros(k, **f32), "v": torch.zeros(k, **f32),
                        "step": torch.zeros(1, dtype=torch.int32, device=dev),
                        "lr": torch.full((1,), self._adaptation_lr, **f32), "scratch": torch.empty(4096, **u8)}

    def _step_fast(self, f, i, finish=True):
        if not finish:
            raise RuntimeError("PPORMA: the data-parallel step is captured as one graph (SNN_DP_GRAPH=single)")
        st = self._fast_state
        mb = st["mb"]
        sl = slice(i * mb, (i + 1) * mb)
        ac, opt = self.actor_critic, self.optimizer
        dev = opt.flat_params.device
        em, es, ws, bs, dw, db = st["actor_tensors"]
        O = st["O"]
        x, x_lat = st["x"], st["x"][:, O:]
        enc, ad, ce, eng = st["enc"], st["ad"], st["critic"], ac.actor.engine
        priv, hist = f["critic_obs"][sl], f["hist"][sl]       # RolloutStorageRMA: privileged obs / observation histories
        side = st.get("side_stream")
        main = torch.cuda.current_stream(dev)
        L = _lib.lib()
        opt.zero_grad()
        # latent = env_factor_encoder(privileged_obs), written straight into the latent columns of cat(obs, latent)
        x[:, :O].copy_(f["obs"][sl])
        enc.repack()
        enc.forward(priv, out_value=x_lat, out_trace=st["enc_trace"])
        if side is not None:
            side.wait_stream(main)
            with torch.cuda.stream(side):
                ce.repack()
                value, ctrace = ce.forward(x, out_value=st["value"], out_trace=st["critic_trace"])
        else:
            ce.repack()
            value, ctrace = ce.forward(x, out_value=st["value"], out_trace=st["critic_trace"])
        mu, trace = eng.forward(x, em.data, es.data, [w.data for w in ws], [b.data for b in bs], dw.data, db.data,
                                train=True, out_mu=st["mu"], out_trace=st["trace"], assume_packed=True)
        if side is not None:
            main.wait_stream(side)
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        with torch.cuda.device(dev):
            _lib.check(L.snn_ppo_head(
                _p(mu), _p(ac.actor.log_std.data), _p(value), _p(f["actions"][sl]), _p(f["logp"][sl]),
                _p(f["adv"][sl]), _p(f["returns"][sl]), _p(f["values"][sl]), _p(f["mu"][sl]), _p(f["sigma"][sl]),
                mb, ac.actor.act_dim, float(self.clip_param), float(self.value_loss_coef), float(self.entropy_coef),
                1 if self.use_clipped_value_loss else 0, _p(st["g_mu"]), _p(st["g_value"]), _p(ac.actor.log_std.grad),
                _p(self._stats), _p(st["scratch"]), st["scratch"].numel(), stream), "snn_ppo_head")
        opt.flat_grads[opt.n:].copy_(self._stats[2:3])
        # both consumers of cat(obs, latent) return the gradient into their input
        if side is not None:
            side.wait_stream(main)
            with torch.cuda.stream(side):
                ce.backward(st["g_value"], ctrace, mb, out=st["critic_grads"], ws=st["critic_ws"], d_x=st["dx_critic"])
        else:
            ce.backward(st["g_value"], ctrace, mb, out=st["critic_grads"], ws=st["critic_ws"], d_x=st["dx_critic"])
        eng.backward(x, mu, st["g_mu"], trace, em.data, es.data, [w.data for w in ws], [b.data for b in bs], dw.data,
                     db.data, need_dobs=True, out=st["grads_out"], assume_packed=True, ws=st["ws_bwd"])
        if side is not None:
            main.wait_stream(side)
        # env_factor_encoder backward on d latent = actor part + critic part
        enc.backward(st["dx_actor"][:, O:], st["enc_trace"], mb, out=st["enc_grads"], ws=st["enc_ws"],
                     g2=st["dx_critic"][:, O:])
        self._finish_step()
        # ---- adaptation-module regression onto the (updated, detached) encoder latent, second Adam (reference :200-213)
        a = self._ad
        world = 1 if self.dp is None else self.dp.world_size
        gsl = opt.flat_grads[a["off"]:a["off"] + a["k"]]
        for _ in range(self.num_adaptation_module_substeps):
            ad.repack()
            enc.repack()
            ad.forward(hist, out_value=st["ad_pred"], out_trace=st["ad_trace"])
            enc.forward(priv, out_value=st["ad_tgt"], out_trace=st["enc_trace"])
            with torch.cuda.device(dev):
                _lib.check(L.snn_mse_grad(_p(st["ad_pred"]), _p(st["ad_tgt"]), st["ad_pred"].numel(), _p(st["ad_g"]),
                                          _p(self._adapt_sum), _p(st["ad_scratch"]),
                                          C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "snn_mse_grad")
            ad.backward(st["ad_g"], st["ad_trace"], mb, out=st["ad_grads"], ws=st["ad_ws"])
            if self.dp is not None:
                self.dp.allreduce_sum(gsl)
            g = opt.param_groups[0]
            with torch.cuda.device(dev):
                _lib.check(L.snn_clip_adam(
                    _p(opt.flat_params[a["off"]:]), _p(gsl), _p(a["m"]), _p(a["v"]), a["k"], _p(a["lr"]), _p(a["step"]),
                    float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]), 0.0, 0.0, 1.0 / world, None,
                    _p(a["scratch"]), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "snn_clip_adam(adaptation)")

    def _graph_signature(self, f, mb):
        return super()._graph_signature(f, mb) + (int(self.num_adaptation_module_substeps), float(self._adaptation_lr))

    def _snapshot(self):
        base = super()._snapshot()
        if self._ad is None:
            return base
        a = self._ad
        return base + (a["m"].clone(), a["v"].clone(), a["step"].clone(), self._adapt_sum.clone())

    def _restore(self, snap):
        if self._ad is not None and len(snap) > 6:
            a = self._ad
            a["m"].copy_(snap[6]); a["v"].copy_(snap[7]); a["step"].copy_(snap[8]); self._adapt_sum.copy_(snap[9])
        super()._restore(snap[:6])

    def refresh_packed(self):
        super().refresh_packed()
        if self._fast_state is not None and self._fast_state.get("enc") is not None:
            self._fast_state["enc"].repack()
            self._fast_state["ad"].repack()

    def init_storage(self, num_envs, num_transitions_per_env, actor_obs_shape, privileged_obs_shape, obs_history_shape,
                     action_shape):
        self.storage = RolloutStorageRMA(num_envs, num_transitions_per_env, actor_obs_shape, privileged_obs_shape,
                                         obs_history_shape, action_shape, self.device)
        if self.dp is not None:
            self.storage.advantage_stats_allreduce = self.dp.allreduce_sum

    def act(self, obs, privileged_obs, obs_history):
        ac = self.actor_critic
        self.transition.actions = ac.act(obs, privileged_obs).detach()
        self.transition.values = ac.evaluate(obs, privileged_obs).detach()
        self.transition.actions_log_prob = ac.get_actions_log_prob(self.transition.actions).detach()
        self.transition.action_mean = ac.action_mean.detach()
        self.transition.action_sigma = ac.action_std.detach()
        self.transition.observations = obs
        self.transition.critic_observations = obs
        self.transition.privileged_observations = privileged_obs
        self.transition.observation_histories = obs_history
        return self.transition.actions

    def compute_returns(self, last_critic_obs, last_critic_privileged_obs):
        last_values = self.actor_critic.evaluate(last_critic_obs, last_critic_privileged_obs).detach()
        self.storage.compute_returns(last_values, self.gamma, self.lam)

    def losses(self, batch):
        (obs_batch, critic_obs_batch, privileged_obs_batch, obs_history_batch, actions_batch, target_values_batch,
         advantages_batch, returns_batch, old_actions_log_prob_batch, old_mu_batch, old_sigma_batch, hid_states_batch,
         masks_batch) = batch
        ac = self.actor_critic
        ac.act(obs_batch, privileged_obs_batch, masks=masks_batch, hidden_states=hid_states_batch[0])
        value_batch = ac.evaluate(critic_obs_batch, privileged_obs_batch, masks=masks_batch,
                                  hidden_states=hid_states_batch[1])
        return self._loss_terms(actions_batch, target_values_batch, advantages_batch, returns_batch,
                                old_actions_log_prob_batch, old_mu_batch, old_sigma_batch, value_batch)

    def update(self):
        num_updates = self.num_learning_epochs * self.num_mini_batches
        if self._fast_supported():
            self._adapt_sum.zero_()
            vl, sl = super().update()                  # the fused steps above, one CUDA graph per minibatch index
            return vl, sl, float(self._adapt_sum.item()) / (num_updates * self.num_adaptation_module_substeps)
        self._stats.zero_()
        self.optimizer.set_lr(self._lr)
        ac = self.actor_critic
        adapt_sum = torch.zeros((), device=self.optimizer.flat_params.device)
        for batch in self.storage.mini_batch_generator(self.num_mini_batches, self.num_learning_epochs):
            self._step_generic(batch)
            privileged_obs_batch, obs_history_batch = batch[2], batch[3]
            for _ in range(self.num_adaptation_module_substeps):
                adaptation_pred = ac.adaptation_module(obs_history_batch)
                with torch.no_grad():
                    adaptation_target = ac.env_factor_encoder(privileged_obs_batch)
                adaptation_loss = torch.nn.functional.mse_loss(adaptation_pred, adaptation_target)
                self.optimizer.zero_grad()             # gradients are views of the flat bucket: zero, never None
                adaptation_loss.backward()
                self.adaptation_optimizer.step()
                adapt_sum += adaptation_loss.detach()
        num_updates = self.num_learning_epochs * self.num_mini_batches
        stats = self._stats.tolist()
        self._lr = self.optimizer.sync_lr_from_device()
        self._params_epoch += 1
        self.storage.clear()
        return (stats[5] / num_updates, stats[4] / num_updates,
                float(adapt_sum) / (num_updates * self.num_adaptation_module_substeps))
