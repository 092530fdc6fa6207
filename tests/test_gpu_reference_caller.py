"""The plugin contract of SURVEY.md 8b, exercised by the reference itself: the UNMODIFIED `PPO` and `RolloutStorage` of
the reference (staged under oracle/_ref by oracle/make_ref.sh; CAS-rl/algorithms/ppo.py:38-187,
CAS-rl/storage/rollout_storage.py:36-184) drive THIS repo's ActorCritic on the GPU — constructor, .to(device),
.parameters() into torch.optim.Adam, act / evaluate / get_actions_log_prob / action_mean / action_std / entropy,
loss.backward() through the sm_100a BPTT kernels, clip_grad_norm_, optimizer.step — and must land on the parameters
the reference's own ActorCritic reaches (ppo_cassie_tiny.npz, written from the unmodified reference)."""
import os

import numpy as np
import pytest
import torch

from cases import PPO_CASE, make_ppo_rollout, make_ppo_state_dict
from helpers import GOLD, rel_err
from oracle import ref_loader as R

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not R.available("cassie"), reason="reference not staged (run oracle/make_ref.sh)")]


def _setup():
    from fullysnnforleggedrobots_b200 import ActorCriticCassie
    pc = PPO_CASE
    g = dict(np.load(os.path.join(GOLD, f"{pc.name}.npz")))
    _, algs, sto = R.load_all(pc.variant)
    ac = ActorCriticCassie(pc.obs_dim, pc.obs_dim, pc.act_dim, actor_hidden_dims=list(pc.hidden),
                           critic_hidden_dims=list(pc.critic_hidden), activation="elu", init_noise_std=1.0)
    ac.load_state_dict(make_ppo_state_dict(pc))
    alg = algs.PPO(ac, num_learning_epochs=pc.epochs, num_mini_batches=1, clip_param=0.2, gamma=pc.gamma, lam=pc.lam,
                   value_loss_coef=1.0, entropy_coef=pc.entropy_coef, learning_rate=pc.lr, max_grad_norm=1.0,
                   use_clipped_value_loss=True, schedule="adaptive", desired_kl=0.01, device="cuda")
    alg.init_storage(pc.num_envs, pc.num_steps, [pc.obs_dim], [None], [pc.act_dim])
    assert type(alg.storage).__module__.startswith("rsl_rl.") and type(alg).__module__.startswith("rsl_rl.")
    return pc, g, ac, alg


def test_reference_ppo_rollout_through_our_actor_critic():
    """Reference PPO.act / process_env_step / compute_returns (ppo.py:90-118) on our module: mu, sigma, values and the
    log-prob of the drawn action match the fixture's policy; GAE matches the fixture when fed the fixture's rollout."""
    pc, g, ac, alg = _setup()
    ro = make_ppo_rollout(pc)
    with torch.inference_mode():
        for s in range(pc.num_steps):
            obs = ro["obs"][s].cuda()
            a = alg.act(obs, obs)
            assert a.shape == (pc.num_envs, pc.act_dim) and a.is_cuda
            assert rel_err(alg.transition.action_mean.cpu(), g["mu"][s]) < 1e-4
            assert rel_err(alg.transition.action_sigma.cpu(), g["sigma"][s]) < 1e-6
            assert rel_err(alg.transition.values.cpu(), g["values"][s]) < 1e-5
            z = (a - alg.transition.action_mean) / alg.transition.action_sigma
            lp = (-0.5 * z * z - torch.log(alg.transition.action_sigma) - 0.9189385332).sum(-1)
            assert rel_err(alg.transition.actions_log_prob.cpu(), lp.cpu()) < 1e-4
            alg.process_env_step(ro["rewards"][s].cuda(), ro["dones"][s].cuda(), {})
        assert alg.storage.step == pc.num_steps
        alg.compute_returns(ro["last_obs"].cuda())
    assert torch.isfinite(alg.storage.advantages).all()


def test_reference_ppo_update_through_our_actor_critic():
    """Reference PPO.update (ppo.py:120-187) with our ActorCritic dropped in == the reference with its own."""
    pc, g, ac, alg = _setup()
    ro, st = make_ppo_rollout(pc), alg.storage
    st.observations.copy_(ro["obs"]); st.actions.copy_(torch.from_numpy(g["actions"]))
    st.rewards.copy_(ro["rewards"].unsqueeze(-1)); st.dones.copy_(ro["dones"].unsqueeze(-1))
    st.values.copy_(torch.from_numpy(g["values"])); st.actions_log_prob.copy_(torch.from_numpy(g["actions_log_prob"]))
    st.mu.copy_(torch.from_numpy(g["mu"])); st.sigma.copy_(torch.from_numpy(g["sigma"]))
    st.step = pc.num_steps
    alg.compute_returns(ro["last_obs"].cuda())
    assert rel_err(st.returns.cpu(), g["returns"]) < 1e-5
    assert rel_err(st.advantages.cpu(), g["advantages"]) < 1e-4
    vl, sl = alg.update()
    assert abs(vl / g["mean_value_loss"] - 1) < 1e-4
    assert abs(sl - g["mean_surrogate_loss"]) < 1e-4 * max(1.0, abs(g["mean_surrogate_loss"]))
    assert abs(alg.learning_rate / g["final_lr"] - 1) < 1e-6
    sd0 = make_ppo_state_dict(pc)
    for k, v in ac.state_dict().items():
        ref = torch.from_numpy(g["final." + k])
        if k == "std":
            assert torch.equal(v.cpu(), ref)          # dead parameter: never receives a gradient, Adam skips it
            continue
        d_ref, d = ref - sd0[k], v.cpu() - sd0[k]
        frac_bad = ((d - d_ref).abs() > 0.02 * pc.lr).float().mean().item()
        assert frac_bad < 0.005, (k, frac_bad)
