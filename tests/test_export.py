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
    assert set(d["optimizer_state_dict"].keys()) == {"state", "param_groups"}         # torch.optim.Adam's layout
    st0 = d["optimizer_state_dict"]["state"][0]
    assert {"step", "exp_avg", "exp_avg_sq"} <= set(st0.keys())
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
