"""Golden fixture of the AMP fork's update: the UNMODIFIED reference AMPPPO.update() (amp_ppo.py:153-274) with the
reference AMPDiscriminator, ReplayBuffer and Normalizer on a seeded rollout, CPU.  Build container only
(/root/reference is not on the GPU box).  Usage:  python tests/golden/make_golden_amp.py

The two random batch sources are replaced by fixed index rules (cases.ExpertPool for the mocap loader, which
needs pybullet_utils; cases.deterministic_replay_generator for ReplayBuffer's np.random.choice) so that the
CUDA implementation can be fed the same rows.
"""
from __future__ import annotations

import contextlib
import importlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from cases import (AMP_CASE, AMP_MIN_STD, AMP_OBS_DIM, AMP_DISC_HIDDEN, AMP_REWARD_COEF, ExpertPool,  # noqa: E402
                   deterministic_replay_generator, make_amp_inputs, make_ppo_rollout, make_ppo_state_dict)
from make_golden import import_ref  # noqa: E402


def main(check=False):
    pc = AMP_CASE
    import_ref("amp", "modules.actor_critic")
    mods = importlib.import_module("rsl_rl.modules")
    amp_ppo = importlib.import_module("rsl_rl.algorithms.amp_ppo")
    disc_mod = importlib.import_module("rsl_rl.algorithms.amp_discriminator")
    utils = importlib.import_module("rsl_rl.utils.utils")
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        ac = mods.ActorCritic(pc.obs_dim, pc.obs_dim, pc.act_dim, actor_hidden_dims=list(pc.hidden),
                              critic_hidden_dims=list(pc.critic_hidden), activation="elu", init_noise_std=1.0)
    ac.actor.encoder.device = ac.actor.snn.device = "cpu"
    ac.load_state_dict(make_ppo_state_dict(pc))
    amp_in = make_amp_inputs(pc)
    disc = disc_mod.AMPDiscriminator(2 * AMP_OBS_DIM, AMP_REWARD_COEF, list(AMP_DISC_HIDDEN), "cpu")
    disc.load_state_dict(amp_in["disc_state_dict"])
    norm = utils.Normalizer(AMP_OBS_DIM)
    norm.update(amp_in["expert_state"].numpy())
    expert = ExpertPool(amp_in["expert_state"], amp_in["expert_next"])
    alg = amp_ppo.AMPPPO(ac, disc, expert, norm, num_learning_epochs=pc.epochs, num_mini_batches=1, clip_param=0.2,
                         gamma=pc.gamma, lam=pc.lam, value_loss_coef=1.0, entropy_coef=pc.entropy_coef,
                         learning_rate=pc.lr, max_grad_norm=1.0, use_clipped_value_loss=True, schedule="adaptive",
                         desired_kl=0.01, device="cpu", amp_replay_buffer_size=24, min_std=AMP_MIN_STD)
    alg.init_storage(pc.num_envs, pc.num_steps, [pc.obs_dim], [None], [pc.act_dim])
    ro = make_ppo_rollout(pc)
    st = alg.storage
    out = {}
    with contextlib.redirect_stdout(io.StringIO()), torch.inference_mode():
        for s in range(pc.num_steps):
            mu = ac.act_inference(ro["obs"][s])
            sigma = torch.exp(ac.actor.log_std).expand_as(mu)
            actions = mu + sigma * ro["noise"][s]
            ac.update_distribution(ro["obs"][s])
            st.observations[s].copy_(ro["obs"][s])
            st.actions[s].copy_(actions)
            st.rewards[s].copy_(ro["rewards"][s].view(-1, 1))
            st.dones[s].copy_(ro["dones"][s].view(-1, 1))
            st.values[s].copy_(ac.evaluate(ro["obs"][s]))
            st.actions_log_prob[s].copy_(ac.get_actions_log_prob(actions).view(-1, 1))
            st.mu[s].copy_(ac.action_mean)
            st.sigma[s].copy_(ac.action_std)
            st.step += 1
            # 8 envs x 4 steps = 32 rows into a 24-row ring: exercises the wrap-around of ReplayBuffer.insert
            alg.amp_storage.insert(amp_in["amp_obs"][s], amp_in["amp_obs"][s + 1])
        alg.compute_returns(ro["last_obs"])
        r, d = disc.predict_amp_reward(amp_in["amp_obs"][0], amp_in["amp_obs"][1], ro["rewards"][0], normalizer=norm)
    out["amp_reward"], out["amp_d"] = r.numpy().copy(), d.numpy().copy()
    out["replay_states"] = alg.amp_storage.states.numpy().copy()
    out["replay_next_states"] = alg.amp_storage.next_states.numpy().copy()
    out["replay_step"], out["replay_num_samples"] = np.int64(alg.amp_storage.step), np.int64(alg.amp_storage.num_samples)
    for k in ("actions", "values", "actions_log_prob", "mu", "sigma", "returns", "advantages"):
        out[k] = getattr(st, k).numpy().copy()
    alg.amp_storage.feed_forward_generator = deterministic_replay_generator(alg.amp_storage)
    with contextlib.redirect_stdout(io.StringIO()):
        res = alg.update()
    out["update_result"] = np.asarray(res, dtype=np.float64)
    out["final_lr"] = np.float64(alg.learning_rate)
    for k, v in ac.state_dict().items():
        out["final." + k] = v.detach().numpy().copy()
    for k, v in disc.state_dict().items():
        out["disc_final." + k] = v.detach().numpy().copy()
    out["norm_mean"], out["norm_var"], out["norm_count"] = norm.mean.copy(), norm.var.copy(), np.float64(norm.count)
    path = os.path.join(HERE, f"{pc.name}.npz")
    from make_golden import _emit          # writes, or with --check compares a fresh reference run with the file
    print("amp case", _emit(path, out, check), "; update() =", res, "lr", alg.learning_rate)


if __name__ == "__main__":
    main(check="--check" in sys.argv[1:])
