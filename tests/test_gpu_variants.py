"""The four fork variants of ActorCritic / PPO on the GPU: RMA's privileged-latent path (d loss / d actor-input flows
into env_factor_encoder), HUMANOID's encoder epsilon, CASSIE's tanh, AMP's fixed_std — gradients checked against a
CPU composite of plain PyTorch MLPs and the spiking-actor oracle on the same weights."""
import pytest
import torch

from helpers import rel_err
from oracle import snn_oracle as O

pytestmark = pytest.mark.gpu


def cpu_twin_grads(ac, actor_in_fn, critic_in_fn, G, H, enc_eps=0.0, dec_tanh=False):
    """loss = sum(mu * G) + sum(value * H) on CPU with identical weights; returns {param name: grad}."""
    sd = {k: v.detach().cpu().clone() for k, v in ac.state_dict().items()}
    leaves = {k: v.requires_grad_(True) for k, v in sd.items() if v.is_floating_point()}

    def mlp(prefix, x):
        i = 0
        while f"{prefix}.{i}.weight" in leaves:
            x = torch.addmm(leaves[f"{prefix}.{i}.bias"], x, leaves[f"{prefix}.{i}.weight"].t())
            if f"{prefix}.{i + 2}.weight" in leaves:
                x = torch.nn.functional.elu(x)
            i += 2
        return x

    p = O.ActorParams.from_state_dict(leaves, spike_ts=5, enc_eps=enc_eps, dec_tanh=dec_tanh)
    # from_state_dict detaches: rebuild with the leaves themselves
    ws, bs, i = [], [], 0
    while f"actor.snn.hidden_layers.{i}.weight" in leaves:
        ws.append(leaves[f"actor.snn.hidden_layers.{i}.weight"]); bs.append(leaves[f"actor.snn.hidden_layers.{i}.bias"]); i += 1
    ws.append(leaves["actor.snn.out_pop_layer.weight"]); bs.append(leaves["actor.snn.out_pop_layer.bias"])
    p = O.ActorParams(leaves["actor.encoder.mean"], leaves["actor.encoder.std"], ws, bs,
                      leaves["actor.decoder.decoder.weight"], leaves["actor.decoder.decoder.bias"], leaves["actor.log_std"],
                      spike_ts=5, enc_eps=enc_eps, dec_tanh=dec_tanh)
    mu, _ = O.actor_forward_autograd(actor_in_fn(mlp), p)
    value = mlp("critic", critic_in_fn(mlp))
    ((mu * G).sum() + (value * H).sum()).backward()
    return mu.detach(), value.detach(), {k: v.grad for k, v in leaves.items() if v.grad is not None}


def test_rma_latent_gradient_path():
    from fullysnnforleggedrobots_b200 import ActorCriticRMA
    torch.manual_seed(11)
    B, no, npriv, nh, na = 96, 20, 9, 30, 4
    ac = ActorCriticRMA(no, npriv, nh, na, actor_hidden_dims=[64, 32], critic_hidden_dims=[32, 16],
                        encoder_hidden_dims=[24, 12], adaptation_hidden_dims=[24, 12], encoder_latent_dims=6).cuda()
    obs, priv = torch.randn(B, no), torch.randn(B, npriv)
    G, H = torch.randn(B, na), torch.randn(B, 1)
    ac.update_distribution(obs.cuda(), priv.cuda())
    mu = ac.action_mean
    value = ac.evaluate(obs.cuda(), priv.cuda())
    ((mu * G.cuda()).sum() + (value * H.cuda()).sum()).backward()
    lat = lambda mlp: mlp("env_factor_encoder", priv)
    mu_ref, v_ref, gr = cpu_twin_grads(ac, lambda mlp: torch.cat((obs, lat(mlp)), -1),
                                       lambda mlp: torch.cat((obs, lat(mlp)), -1), G, H)
    assert rel_err(mu.detach().cpu(), mu_ref) < 1e-4
    assert rel_err(value.detach().cpu(), v_ref) < 1e-4
    for name, prm in ac.named_parameters():
        if name.startswith("encoder."):        # alias of env_factor_encoder (registered twice, like the reference)
            continue
        if name in gr:
            assert rel_err(prm.grad.cpu(), gr[name]) < 2e-4, name
    assert ac.env_factor_encoder[0].weight.grad.abs().max() > 0        # the latent path carries gradient
    assert ac.adaptation_module[0].weight.grad is None
    pol = {}
    a = ac.act_inference(obs.cuda(), torch.randn(B, nh).cuda(), policy_info=pol)
    assert a.shape == (B, na) and "latents" in pol


def test_rma_ppo_update_runs_and_trains_adaptation_module():
    from fullysnnforleggedrobots_b200 import PPORMA, ActorCriticRMA
    torch.manual_seed(12)
    N, S, no, npriv, nh, na = 64, 6, 20, 9, 30, 4
    ac = ActorCriticRMA(no, npriv, nh, na, actor_hidden_dims=[64, 32], critic_hidden_dims=[32, 16],
                        encoder_hidden_dims=[24, 12], adaptation_hidden_dims=[24, 12], encoder_latent_dims=6)
    alg = PPORMA(ac, num_learning_epochs=2, num_mini_batches=2, entropy_coef=0.01, schedule="adaptive", device="cuda",
                 num_adaptation_module_substeps=2)
    alg.init_storage(N, S, [no], [npriv], [nh], [na])
    w0 = ac.adaptation_module[0].weight.detach().clone()
    e0 = ac.env_factor_encoder[0].weight.detach().clone()
    with torch.inference_mode():
        for _ in range(S):
            obs, priv, hist = torch.randn(N, no).cuda(), torch.randn(N, npriv).cuda(), torch.randn(N, nh).cuda()
            alg.act(obs, priv, hist)
            alg.process_env_step(torch.randn(N).cuda(), (torch.rand(N) < 0.1).cuda(), {})
        alg.compute_returns(obs, priv)
    vl, sl, al = alg.update()
    assert all(x == x for x in (vl, sl, al)) and al > 0
    assert not torch.equal(ac.adaptation_module[0].weight, w0)
    assert not torch.equal(ac.env_factor_encoder[0].weight, e0)
    assert all(torch.isfinite(p).all() for p in ac.parameters())


@pytest.mark.parametrize("mode", ["default", "generic"])
def test_rma_ppo_update_matches_reference_fixture(mode):
    """The RMA fork's whole PPO.update (RMA-rl/algorithms/ppo.py:134-221: PPO step through the privileged-latent path +
    adaptation-module regression substeps with a second Adam) against rma_ppo_tiny.npz, written by
    tests/golden/make_golden.py from the unmodified reference.  mode "default" = whatever PPORMA picks (the fused /
    graphed path where supported), "generic" = the op-by-op autograd path."""
    import os

    import numpy as np

    from cases import RMA_CASE as rc, make_rma_rollout, make_rma_state_dict
    from fullysnnforleggedrobots_b200 import PPORMA, ActorCriticRMA
    from helpers import GOLD
    g = dict(np.load(os.path.join(GOLD, f"{rc.name}.npz")))
    ro = make_rma_rollout(rc)
    ac = ActorCriticRMA(rc.num_obs, rc.num_priv, rc.num_hist, rc.act_dim, actor_hidden_dims=list(rc.hidden),
                        critic_hidden_dims=list(rc.critic_hidden), encoder_hidden_dims=list(rc.enc_hidden),
                        adaptation_hidden_dims=list(rc.adapt_hidden), encoder_latent_dims=rc.latent)
    sd0 = make_rma_state_dict(rc)
    ac.load_state_dict(sd0)
    kw = {} if mode == "default" else {"fused": False}
    alg = PPORMA(ac, num_learning_epochs=rc.epochs, num_mini_batches=1, clip_param=0.2, gamma=rc.gamma, lam=rc.lam,
                 value_loss_coef=1.0, entropy_coef=rc.entropy_coef, learning_rate=rc.lr, max_grad_norm=1.0,
                 use_clipped_value_loss=True, schedule="adaptive", desired_kl=0.01, device="cuda",
                 num_adaptation_module_substeps=rc.substeps, **kw)
    alg.init_storage(rc.num_envs, rc.num_steps, [rc.num_obs], [rc.num_priv], [rc.num_hist], [rc.act_dim])
    # rollout quantities: act() on the fixture's observations reproduces the reference's mu / values
    with torch.inference_mode():
        alg.act(ro["obs"][0].cuda(), ro["priv"][0].cuda(), ro["hist"][0].cuda())
    assert rel_err(alg.transition.action_mean.cpu(), g["mu"][0]) < 1e-4
    assert rel_err(alg.transition.values.cpu(), g["values"][0]) < 1e-5
    alg.transition.clear()
    st = alg.storage
    st.observations.copy_(ro["obs"]); st.privileged_observations.copy_(ro["priv"]); st.observation_histories.copy_(ro["hist"])
    st.actions.copy_(torch.from_numpy(g["actions"])); st.rewards.copy_(ro["rewards"].unsqueeze(-1))
    st.dones.copy_(ro["dones"].unsqueeze(-1)); st.values.copy_(torch.from_numpy(g["values"]))
    st.actions_log_prob.copy_(torch.from_numpy(g["actions_log_prob"]))
    st.mu.copy_(torch.from_numpy(g["mu"])); st.sigma.copy_(torch.from_numpy(g["sigma"]))
    st.step = rc.num_steps
    alg.compute_returns(ro["last_obs"].cuda(), ro["last_priv"].cuda())
    assert rel_err(st.returns.cpu(), g["returns"]) < 1e-5
    assert rel_err(st.advantages.cpu(), g["advantages"]) < 1e-4
    vl, sl, al = alg.update()
    if mode == "default":          # the fused step (MLPs on the tensor-core kernels), replayed as one CUDA graph
        assert alg._fast_supported() and len(alg._graphs) == 1 and alg._fast_state.get("enc") is not None
    else:
        assert not alg._fast_supported() and len(alg._graphs) == 0
    assert abs(vl / g["mean_value_loss"] - 1) < 1e-4
    assert abs(sl - g["mean_surrogate_loss"]) < 1e-4 * max(1.0, abs(g["mean_surrogate_loss"]))
    assert abs(al / g["mean_adaptation_loss"] - 1) < 1e-3
    assert abs(alg.learning_rate / g["final_lr"] - 1) < 1e-6
    for k, v in ac.state_dict().items():
        ref = torch.from_numpy(g["final." + k])
        if k == "std":
            assert torch.equal(v.cpu(), ref)
            continue
        d_ref, d = ref - sd0[k], v.cpu() - sd0[k]
        frac_bad = ((d - d_ref).abs() > 0.02 * rc.lr).float().mean().item()
        assert frac_bad < 0.005, (k, frac_bad)
        if k.startswith(("adaptation_module", "env_factor_encoder")):
            assert d_ref.abs().max() > 0 and d.abs().max() > 0          # both modules really trained


@pytest.mark.parametrize("variant,enc_eps,dec_tanh", [("amp", 0.0, False), ("humanoid", 1e-5, False), ("cassie", 0.0, True)])
def test_fork_variants_match_cpu_twin(variant, enc_eps, dec_tanh):
    from fullysnnforleggedrobots_b200 import VARIANTS
    torch.manual_seed(13)
    B, no, na = 80, 14, 3
    ac = VARIANTS[variant](no, no, na, actor_hidden_dims=[48, 24], critic_hidden_dims=[32, 16]).cuda()
    obs = torch.randn(B, no)
    G, H = torch.randn(B, na), torch.randn(B, 1)
    ac.update_distribution(obs.cuda())
    mu, value = ac.action_mean, ac.evaluate(obs.cuda())
    ((mu * G.cuda()).sum() + (value * H.cuda()).sum()).backward()
    mu_ref, v_ref, gr = cpu_twin_grads(ac, lambda mlp: obs, lambda mlp: obs, G, H, enc_eps=enc_eps, dec_tanh=dec_tanh)
    assert rel_err(mu.detach().cpu(), mu_ref) < 1e-4
    for name, prm in ac.named_parameters():
        if name in gr:
            assert rel_err(prm.grad.cpu(), gr[name]) < 2e-4, name
    assert ac.std.grad is None                       # dead parameter in SNN mode, as in the reference
    if variant == "amp":
        fixed = VARIANTS[variant](no, no, na, actor_hidden_dims=[16], critic_hidden_dims=[16], fixed_std=True)
        assert not isinstance(fixed.std, torch.nn.Parameter) and fixed.fixed_std
