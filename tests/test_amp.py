"""AMP fork: AMPPPO / AMPDiscriminator / ReplayBuffer / Normalizer against the fixture written by the UNMODIFIED
reference (tests/golden/make_golden_amp.py -> amp_ppo_tiny.npz; amp_ppo.py:153-274, amp_discriminator.py,
replay_buffer.py, utils.py:79-131).  Host helpers are checked on the CPU, the update on the GPU."""
import os

import numpy as np
import pytest
import torch

from cases import (AMP_CASE, AMP_DISC_HIDDEN, AMP_MIN_STD, AMP_OBS_DIM, AMP_REWARD_COEF, ExpertPool,
                   deterministic_replay_generator, make_amp_inputs, make_ppo_rollout, make_ppo_state_dict)
from helpers import GOLD, rel_err


def load():
    return dict(np.load(os.path.join(GOLD, f"{AMP_CASE.name}.npz")))


def test_replay_buffer_ring_matches_reference():
    from fullysnnforleggedrobots_b200.amp import ReplayBuffer
    pc, g, a = AMP_CASE, load(), make_amp_inputs(AMP_CASE)
    buf = ReplayBuffer(AMP_OBS_DIM, 24, "cpu")
    for s in range(pc.num_steps):
        buf.insert(a["amp_obs"][s], a["amp_obs"][s + 1])
    assert buf.step == int(g["replay_step"]) and buf.num_samples == int(g["replay_num_samples"])
    assert np.array_equal(buf.states.numpy(), g["replay_states"])
    assert np.array_equal(buf.next_states.numpy(), g["replay_next_states"])
    s, n = next(buf.feed_forward_generator(1, 40))
    assert s.shape == (40, AMP_OBS_DIM) and n.shape == (40, AMP_OBS_DIM)


def test_normalizer_and_amp_reward_match_reference():
    from fullysnnforleggedrobots_b200.amp import AMPDiscriminator, Normalizer
    pc, g, a, ro = AMP_CASE, load(), make_amp_inputs(AMP_CASE), make_ppo_rollout(AMP_CASE)
    norm = Normalizer(AMP_OBS_DIM)
    norm.update(a["expert_state"].numpy())                # numpy, as the reference's callers pass it
    disc = AMPDiscriminator(2 * AMP_OBS_DIM, AMP_REWARD_COEF, list(AMP_DISC_HIDDEN), "cpu")
    disc.load_state_dict(a["disc_state_dict"])
    r, d = disc.predict_amp_reward(a["amp_obs"][0], a["amp_obs"][1], ro["rewards"][0], normalizer=norm)
    assert rel_err(r, g["amp_reward"]) < 1e-5 and rel_err(d, g["amp_d"]) < 1e-5
    x = a["amp_obs"][0].numpy()
    assert np.allclose(norm.normalize(x), norm.normalize_torch(a["amp_obs"][0]).numpy(), atol=1e-6)
    assert sorted(disc.state_dict().keys()) == sorted(k[len("disc_final."):] for k in g if k.startswith("disc_final."))


def make_alg(pc, a, **kw):
    from fullysnnforleggedrobots_b200 import AMPPPO, ActorCritic
    from fullysnnforleggedrobots_b200.amp import AMPDiscriminator, Normalizer
    ac = ActorCritic(pc.obs_dim, pc.obs_dim, pc.act_dim, actor_hidden_dims=list(pc.hidden),
                     critic_hidden_dims=list(pc.critic_hidden), activation="elu", init_noise_std=1.0)
    ac.load_state_dict(make_ppo_state_dict(pc))
    disc = AMPDiscriminator(2 * AMP_OBS_DIM, AMP_REWARD_COEF, list(AMP_DISC_HIDDEN), "cuda")
    disc.load_state_dict(a["disc_state_dict"])
    norm = Normalizer(AMP_OBS_DIM)
    norm.update(a["expert_state"])
    expert = ExpertPool(a["expert_state"].cuda(), a["expert_next"].cuda())
    alg = AMPPPO(ac, disc, expert, norm, num_learning_epochs=pc.epochs, num_mini_batches=1, clip_param=0.2,
                 gamma=pc.gamma, lam=pc.lam, value_loss_coef=1.0, entropy_coef=pc.entropy_coef, learning_rate=pc.lr,
                 max_grad_norm=1.0, use_clipped_value_loss=True, schedule="adaptive", desired_kl=0.01, device="cuda",
                 amp_replay_buffer_size=24, min_std=AMP_MIN_STD, **kw)
    alg.init_storage(pc.num_envs, pc.num_steps, [pc.obs_dim], [None], [pc.act_dim])
    return ac, disc, norm, alg


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["fused_graph", "generic", "fused_graph_tf32"])
def test_amp_update_matches_reference(mode):
    """fp32 discriminator GEMMs: the reference fixture's tolerances; tf32 (the default, = the reference's pinned
    torch 1.10 behaviour on tensor-core GPUs): same checks with the discriminator-side tolerances at TF32 level."""
    pc, g, a, ro = AMP_CASE, load(), make_amp_inputs(AMP_CASE), make_ppo_rollout(AMP_CASE)
    tf32 = mode.endswith("tf32")
    ac, disc, norm, alg = make_alg(pc, a, fused=mode != "generic", use_cuda_graph=mode.startswith("fused_graph"),
                                   disc_matmul="tf32" if tf32 else "fp32")
    assert alg.disc_matmul == ("tf32" if tf32 else "fp32")
    # rollout through the public act / process_env_step surface (amp_obs rides along), then overwrite the sampled
    # quantities with the reference's (its noise stream cannot be reproduced)
    with torch.inference_mode():
        for s in range(pc.num_steps):
            obs = ro["obs"][s].cuda()
            alg.act(obs, obs, a["amp_obs"][s].cuda())
            assert rel_err(alg.transition.action_mean.cpu(), g["mu"][s]) < 1e-4
            alg.process_env_step(ro["rewards"][s].cuda(), ro["dones"][s].cuda(), {}, a["amp_obs"][s + 1].cuda())
    assert np.array_equal(alg.amp_storage.states.cpu().numpy(), g["replay_states"])
    st = alg.storage
    for k in ("actions", "values", "actions_log_prob", "mu", "sigma"):
        getattr(st, k).copy_(torch.from_numpy(g[k]))
    alg.compute_returns(ro["last_obs"].cuda())
    assert rel_err(st.advantages.cpu(), g["advantages"]) < 1e-4
    alg.amp_storage.feed_forward_generator = deterministic_replay_generator(alg.amp_storage)
    res = alg.update()
    ref = g["update_result"]
    assert len(res) == 6
    for i, (x, y) in enumerate(zip(res, ref)):
        tol = 5e-3 if (tf32 and i >= 2) else 2e-4           # entries 2.. are discriminator quantities
        assert abs(x - y) <= tol * max(1.0, abs(y)), (i, x, y)
    assert abs(alg.learning_rate / g["final_lr"] - 1) < 1e-6
    assert torch.equal(ac.std.detach().cpu(), torch.full((pc.act_dim,), AMP_MIN_STD))     # min_std clamp (:251-252)
    sd0 = make_ppo_state_dict(pc)
    for k, v in ac.state_dict().items():
        if k == "std":
            continue
        d_ref = torch.from_numpy(g["final." + k]) - sd0[k]
        d = v.cpu() - sd0[k]
        frac_bad = ((d - d_ref).abs() > 0.02 * pc.lr).float().mean().item()
        assert frac_bad < 0.005, (k, frac_bad)
    # discriminator: same Adam (weight-decayed groups, adaptive learning rate shared with the actor-critic)
    for k, v in disc.state_dict().items():
        d_ref = torch.from_numpy(g["disc_final." + k]) - a["disc_state_dict"][k]
        d = v.cpu() - a["disc_state_dict"][k]
        frac_bad = ((d - d_ref).abs() > (0.2 if tf32 else 0.02) * pc.lr).float().mean().item()
        assert frac_bad < (0.02 if tf32 else 0.005), (k, frac_bad)
    assert np.allclose(norm.mean.cpu().numpy(), g["norm_mean"], rtol=1e-5, atol=1e-6)
    assert np.allclose(norm.var.cpu().numpy(), g["norm_var"], rtol=1e-5, atol=1e-6)
    assert abs(norm.count - float(g["norm_count"])) < 1e-9


@pytest.mark.gpu
def test_amp_discriminator_graph_replay_matches_eager():
    """The discriminator step is captured into a CUDA graph on its third call.  Three updates (6 steps: 2 eager
    warm-ups, capture + 4 replays) must leave exactly what the all-eager path leaves."""
    pc, g, a, ro = AMP_CASE, load(), make_amp_inputs(AMP_CASE), make_ppo_rollout(AMP_CASE)

    def run(disc_graph):
        ac, disc, norm, alg = make_alg(pc, a)
        alg.disc_cuda_graph = disc_graph
        outs = []
        for _ in range(3):
            with torch.inference_mode():
                for s in range(pc.num_steps):
                    obs = ro["obs"][s].cuda()
                    alg.act(obs, obs, a["amp_obs"][s].cuda())
                    alg.process_env_step(ro["rewards"][s].cuda(), ro["dones"][s].cuda(), {}, a["amp_obs"][s + 1].cuda())
            st = alg.storage
            for k in ("actions", "values", "actions_log_prob", "mu", "sigma"):
                getattr(st, k).copy_(torch.from_numpy(g[k]))
            alg.compute_returns(ro["last_obs"].cuda())
            alg.amp_storage.feed_forward_generator = deterministic_replay_generator(alg.amp_storage)
            outs.append(alg.update())
        graphed = alg._disc_state is not None and alg._disc_state["graph"] is not None
        return outs, {k: v.detach().cpu().clone() for k, v in disc.state_dict().items()}, norm, graphed

    o_e, sd_e, n_e, g_e = run(False)
    o_g, sd_g, n_g, g_g = run(True)
    assert g_g and not g_e
    for x, y in zip(o_e, o_g):
        assert np.allclose(x, y, rtol=1e-5, atol=1e-6), (x, y)
    for k in sd_e:
        assert rel_err(sd_g[k], sd_e[k]) < 1e-5, k
    assert np.allclose(n_g.mean.cpu().numpy(), n_e.mean.cpu().numpy(), rtol=1e-9)
    assert abs(float(n_g.count) - float(n_e.count)) < 1e-9
