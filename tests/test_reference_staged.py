"""The staged copy of the reference (oracle/_ref, written by oracle/make_ref.sh from /root/reference) is what the GPU
box uses as the caller (tests/test_gpu_reference_caller.py) and as the CPU arm of `bench.py --impl reference`.
CPU-only: it imports, it is byte-identical to the tree it was copied from (when that tree is present), and its
PopSpikeActor / PPO.update agree with the oracle restatement and with the committed golden fixtures."""
import hashlib
import math
import os

import numpy as np
import pytest
import torch

from cases import PPO_CASE, case_by_name, make_obs, make_ppo_rollout, make_ppo_state_dict, make_state_dict
from helpers import GOLD, load_golden, oracle_params, rel_err
from oracle import ref_loader as R
from oracle import snn_oracle as O

needs_ref = pytest.mark.skipif(not R.available("cassie"), reason="reference not staged (run oracle/make_ref.sh)")


@needs_ref
def test_staged_files_are_unmodified_copies():
    if R.kind("cassie") != "staged" or not os.path.isdir(R.REFERENCE):
        pytest.skip("needs both oracle/_ref and /root/reference (build container)")
    n = 0
    for fork, rel in R.FORK_ROOT.items():
        for dirpath, _, files in os.walk(os.path.join(R.STAGED, fork)):
            for f in files:
                if not f.endswith(".py"):
                    continue
                staged = os.path.join(dirpath, f)
                src = os.path.join(R.REFERENCE, rel, os.path.relpath(staged, os.path.join(R.STAGED, fork)))
                assert hashlib.sha256(open(staged, "rb").read()).digest() == hashlib.sha256(open(src, "rb").read()).digest(), staged
                n += 1
    assert n >= 40


@needs_ref
@pytest.mark.parametrize("name", ["tiny_amp", "tiny_cassie_tanh", "tiny_humanoid_eps", "tiny_rma_T8"])
def test_staged_actor_matches_oracle_and_golden(name):
    c = case_by_name(name)
    mods = R.load(c.variant, "modules.actor_critic")
    with R.quiet():
        actor = mods.PopSpikeActor(c.obs_dim, c.act_dim, c.pop, c.pop, list(c.hidden), (-5, 5), math.sqrt(0.15),
                                   spike_ts=c.spike_ts, device="cpu")
        actor.load_state_dict({k[len("actor."):]: v for k, v in make_state_dict(c).items()})
        mu, std = actor(make_obs(c))
    mu_o, std_o = O.actor_forward(make_obs(c), oracle_params(c))
    assert torch.equal(mu.detach(), mu_o) or rel_err(mu.detach(), mu_o) < 1e-6
    assert rel_err(mu.detach(), load_golden(name)["mu"]) < 1e-6


@needs_ref
def test_staged_ppo_update_reproduces_the_fixture():
    """The reference's own PPO.update (CASSIE copy) from the staged files == ppo_cassie_tiny.npz."""
    pc = PPO_CASE
    g = dict(np.load(os.path.join(GOLD, f"{pc.name}.npz")))
    mods, algs, _ = R.load_all(pc.variant)
    with R.quiet():
        ac = mods.ActorCritic(pc.obs_dim, pc.obs_dim, pc.act_dim, actor_hidden_dims=list(pc.hidden),
                              critic_hidden_dims=list(pc.critic_hidden), activation="elu", init_noise_std=1.0)
    ac.actor.encoder.device = ac.actor.snn.device = "cpu"
    ac.load_state_dict(make_ppo_state_dict(pc))
    alg = algs.PPO(ac, num_learning_epochs=pc.epochs, num_mini_batches=1, clip_param=0.2, gamma=pc.gamma, lam=pc.lam,
                   value_loss_coef=1.0, entropy_coef=pc.entropy_coef, learning_rate=pc.lr, max_grad_norm=1.0,
                   use_clipped_value_loss=True, schedule="adaptive", desired_kl=0.01, device="cpu")
    alg.init_storage(pc.num_envs, pc.num_steps, [pc.obs_dim], [None], [pc.act_dim])
    ro, st = make_ppo_rollout(pc), alg.storage
    st.observations.copy_(ro["obs"]); st.actions.copy_(torch.from_numpy(g["actions"]))
    st.rewards.copy_(ro["rewards"].unsqueeze(-1)); st.dones.copy_(ro["dones"].unsqueeze(-1))
    st.values.copy_(torch.from_numpy(g["values"])); st.actions_log_prob.copy_(torch.from_numpy(g["actions_log_prob"]))
    st.mu.copy_(torch.from_numpy(g["mu"])); st.sigma.copy_(torch.from_numpy(g["sigma"]))
    st.step = pc.num_steps
    with R.quiet():
        alg.compute_returns(ro["last_obs"])
        vl, sl = alg.update()
    assert abs(vl / g["mean_value_loss"] - 1) < 1e-6 and abs(alg.learning_rate / g["final_lr"] - 1) < 1e-9
    for k, v in ac.state_dict().items():
        assert rel_err(v, g["final." + k]) < 1e-6, k
