"""Checkpoint layout of the reference runner (on_policy_runner.py:224-244) and the deployment artefact that replaces
export_policy_as_jit (legged_gym/utils/helpers.py:180-191)."""
import math
import os

import numpy as np
import pytest
import torch

from cases import PPO_CASE, make_ppo_rollout, make_ppo_state_dict
from helpers import GOLD, rel_err


def test_artefact_is_plain_tensors_and_descriptor(tmp_path):
    from fullysnnforleggedrobots_b200 import PopSpikeActor, export_policy, load_policy
    actor = PopSpikeActor(7, 3, 10, 10, [24, 16], (-5, 5), math.sqrt(0.15), spike_ts=5, enc_eps=1e-5)
    path = str(tmp_path / "policy.pt")
    desc = export_policy(actor, path)
    assert desc["hidden"] == [24, 16] and desc["enc_eps"] == pytest.approx(1e-5) and desc["dec_tanh"] is False
    d = torch.load(path, map_location="cpu", weights_only=True)        # nothing pickled
    assert d["format"] == "snn_b200_policy_v1"
    assert sorted(d["state_dict"].keys()) == sorted(actor.state_dict().keys())
    pol = load_policy(path, device="cpu")
    for k, v in actor.state_dict().items():
        assert torch.equal(pol.actor.state_dict()[k], v)
    with pytest.raises(RuntimeError):                                   # no CPU fallback for the spiking path
        pol(torch.zeros(2, 7))
    torch.save({"format": "something else"}, path)
    with pytest.raises(ValueError):
        load_policy(path, device="cpu")


@pytest.mark.gpu
def test_exported_policy_reproduces_act_inference(tmp_path):
    from fullysnnforleggedrobots_b200 import ActorCriticCassie, export_policy, load_policy
    pc = PPO_CASE
    ac = ActorCriticCassie(pc.obs_dim, pc.obs_dim, pc.act_dim, actor_hidden_dims=list(pc.hidden),
                           critic_hidden_dims=list(pc.critic_hidden), activation="elu", init_noise_std=1.0).cuda()
    ac.load_state_dict(make_ppo_state_dict(pc))
    path = str(tmp_path / "policy.pt")
    export_policy(ac, path)
    pol = load_policy(path)
    obs = make_ppo_rollout(pc)["obs"][0].cuda()
    with torch.inference_mode():
        ref = ac.act_inference(obs)
    assert torch.equal(pol(obs), ref)                                   # same kernels, same weights: bitwise
    g = dict(np.load(os.path.join(GOLD, f"{pc.name}.npz")))
    assert rel_err(pol(obs).cpu(), g["mu"][0]) < 1e-4                   # and it is the reference's action mean


@pytest.mark.gpu
def test_reference_checkpoint_layout_round_trip(tmp_path):
    """A `model_*.pt` in the reference runner's layout (here: the reference's own post-update parameters from the
    fixture + a torch.optim.Adam state dict) loads into the fused PPO, and a checkpoint written here loads back."""
    from test_gpu_ppo import fill_storage, make_alg
    from fullysnnforleggedrobots_b200 import load_checkpoint, save_checkpoint
    pc, ro = PPO_CASE, make_ppo_rollout(PPO_CASE)
    g = dict(np.load(os.path.join(GOLD, f"{pc.name}.npz")))
    ac, alg = make_alg(pc)
    fill_storage(alg, pc, g, ro)
    alg.compute_returns(ro["last_obs"].cuda())
    alg.update()
    path = str(tmp_path / "model_1.pt")
    save_checkpoint(path, alg, it=1, infos={"note": "x"})
    d = torch.load(path, map_location="cpu", weights_only=False)
    assert sorted(d.keys()) == ["infos", "iter", "model_state_dict", "optimizer_state_dict"]
    osd = d["optimizer_state_dict"]
    assert set(osd.keys()) == {"state", "param_groups"}                                # torch.optim.Adam's layout ...
    n_ref = len(list(ac.parameters()))
    assert osd["param_groups"][0]["params"] == list(range(n_ref))                      # ... in the REFERENCE's indexing:
    assert 0 not in osd["state"] and set(osd["state"].keys()) == set(range(1, n_ref))  # `std` is parameter 0, stateless
    assert {"step", "exp_avg", "exp_avg_sq"} <= set(osd["state"][1].keys())
    # a stock torch.optim.Adam over actor_critic.parameters() (what the reference runner builds) loads it
    stock = torch.optim.Adam(ac.parameters(), lr=1e-3)
    stock.load_state_dict(osd)
    p1 = list(ac.parameters())[1]
    assert torch.equal(stock.state[p1]["exp_avg"], alg.optimizer.state[p1]["exp_avg"])
    # reference-produced parameters into a fresh algorithm object, then the saved optimizer state
    ac2, alg2 = make_alg(pc)
    ref_sd = {k[len("final."):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("final.")}
    infos = load_checkpoint({"model_state_dict": ref_sd, "optimizer_state_dict": d["optimizer_state_dict"],
                             "iter": 7, "infos": {"a": 1}}, alg2)
    assert infos == {"a": 1}
    for k, v in ac2.state_dict().items():
        assert torch.equal(v.cpu(), ref_sd[k]), k
    assert rel_err(alg2.optimizer.exp_avg.cpu(), alg.optimizer.exp_avg.cpu()) == 0.0
    assert int(alg2.optimizer.step_dev.item()) == int(alg.optimizer.step_dev.item())
    # parameters are still views of the flat bucket and the next fused update uses the loaded weights
    assert ac2.actor.snn.out_pop_layer.weight.data_ptr() >= alg2.optimizer.flat_params.data_ptr()
    obs = ro["obs"][0].cuda()
    with torch.inference_mode():
        mu_loaded = ac2.act_inference(obs)
    fill_storage(alg2, pc, g, ro)
    alg2.compute_returns(ro["last_obs"].cuda())
    vl, sl = alg2.update()
    assert vl == vl and sl == sl
    with torch.inference_mode():
        assert not torch.equal(ac2.act_inference(obs), mu_loaded)
    # our own file loads back bit-exactly
    load_checkpoint(path, alg2)
    for k, v in ac.state_dict().items():
        assert torch.equal(ac2.state_dict()[k], v), k


def test_reference_optimizer_state_is_split_by_index_surgery():
    """CPU: reference-indexed Adam state (dead `std` at index 0, AMP's two discriminator groups behind the
    actor-critic group) -> FlatAdam's indexing + the discriminator optimizer's indexing."""
    from fullysnnforleggedrobots_b200.export import split_reference_optimizer_state
    sd = {"state": {1: {"k": 1}, 2: {"k": 2}, 3: {"k": 3}, 4: {"k": 4}, 6: {"k": 6}},
          "param_groups": [{"lr": 1e-3, "params": [0, 1, 2, 3], "name": "actor_critic"},
                           {"lr": 1e-3, "weight_decay": 1e-3, "params": [4, 5], "name": "amp_trunk"},
                           {"lr": 1e-3, "weight_decay": 1e-1, "params": [6], "name": "amp_head"}]}
    ac, disc = split_reference_optimizer_state(sd, 4, [1, 2, 3], 3)
    assert ac["state"] == {0: {"k": 1}, 1: {"k": 2}, 2: {"k": 3}} and ac["param_groups"][0]["params"] == [0, 1, 2]
    assert disc["state"] == {0: {"k": 4}, 2: {"k": 6}}
    assert [g["params"] for g in disc["param_groups"]] == [[0, 1], [2]]
    assert [g["weight_decay"] for g in disc["param_groups"]] == [1e-3, 1e-1]
    # the std-less layout written by earlier versions of this package passes through
    old = {"state": {0: {"k": 1}}, "param_groups": [{"lr": 1e-3, "params": [0, 1, 2]}]}
    ac2, disc2 = split_reference_optimizer_state(old, 4, [1, 2, 3], 3)
    assert ac2["state"] == {0: {"k": 1}} and disc2 is None
    with pytest.raises(ValueError):
        split_reference_optimizer_state({"state": {}, "param_groups": [{"params": [0, 1]}]}, 4, [1, 2, 3], 3)


@pytest.mark.gpu
def test_amp_checkpoint_round_trip(tmp_path):
    """AMP runner layout (amp_on_policy_runner.py:268-286): one optimizer state with three groups, the discriminator
    and the normaliser; written by one AMPPPO, loaded by a fresh one, which then continues identically."""
    from cases import AMP_CASE, make_amp_inputs, deterministic_replay_generator
    from test_amp import make_alg as make_amp
    from fullysnnforleggedrobots_b200 import load_checkpoint, save_checkpoint
    pc, a, ro = AMP_CASE, make_amp_inputs(AMP_CASE), make_ppo_rollout(AMP_CASE)
    g = dict(np.load(os.path.join(GOLD, f"{pc.name}.npz")))

    def one_update(alg):
        with torch.inference_mode():
            for s in range(pc.num_steps):
                obs = ro["obs"][s].cuda()
                alg.act(obs, obs, a["amp_obs"][s].cuda())
                alg.process_env_step(ro["rewards"][s].cuda(), ro["dones"][s].cuda(), {}, a["amp_obs"][s + 1].cuda())
        for k in ("actions", "values", "actions_log_prob", "mu", "sigma"):
            getattr(alg.storage, k).copy_(torch.from_numpy(g[k]))
        alg.compute_returns(ro["last_obs"].cuda())
        alg.amp_storage.feed_forward_generator = deterministic_replay_generator(alg.amp_storage)
        return alg.update()

    ac, disc, norm, alg = make_amp(pc, a)
    one_update(alg)
    path = str(tmp_path / "model_amp.pt")
    save_checkpoint(path, alg, it=3)
    d = torch.load(path, map_location="cpu", weights_only=False)
    assert [gr["name"] for gr in d["optimizer_state_dict"]["param_groups"]] == ["actor_critic", "amp_trunk", "amp_head"]
    assert "discriminator_state_dict" in d and set(d["amp_normalizer"].keys()) == {"mean", "var", "count"}
    ac2, disc2, norm2, alg2 = make_amp(pc, a)
    load_checkpoint(path, alg2)
    for k, v in disc.state_dict().items():
        assert torch.equal(disc2.state_dict()[k], v), k
    assert torch.equal(norm2.mean.cpu(), norm.mean.cpu()) and float(norm2.count) == float(norm.count)
    assert abs(alg2.learning_rate / alg.learning_rate - 1) < 1e-6
    # the AMP replay ring is not part of a checkpoint (nor in the reference): give both the same one
    for k in ("states", "next_states"):
        getattr(alg2.amp_storage, k).copy_(getattr(alg.amp_storage, k))
    alg2.amp_storage.step, alg2.amp_storage.num_samples = alg.amp_storage.step, alg.amp_storage.num_samples
    r1, r2 = one_update(alg), one_update(alg2)                # same state -> same next update
    assert np.allclose(r1, r2, rtol=1e-5, atol=1e-6), (r1, r2)
    for k, v in disc.state_dict().items():
        assert rel_err(disc2.state_dict()[k], v) < 1e-5, k
    for k, v in ac.state_dict().items():
        assert rel_err(ac2.state_dict()[k], v) < 1e-5, k
