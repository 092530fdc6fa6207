"""GPU parity of the actor-critic / PPO / rollout-storage surface against the reference fixture
(ppo_cassie_tiny.npz: the reference's own PPO.update on a seeded rollout) and the CPU oracle."""
import os

import numpy as np
import pytest
import torch

from cases import PPO_CASE, make_ppo_rollout, make_ppo_state_dict
from helpers import GOLD, rel_err

pytestmark = pytest.mark.gpu


def load():
    return dict(np.load(os.path.join(GOLD, f"{PPO_CASE.name}.npz")))


def make_alg(pc, **kw):
    from fullysnnforleggedrobots_b200 import PPO, ActorCriticCassie
    ac = ActorCriticCassie(pc.obs_dim, pc.obs_dim, pc.act_dim, actor_hidden_dims=list(pc.hidden),
                           critic_hidden_dims=list(pc.critic_hidden), activation="elu", init_noise_std=1.0)
    ac.load_state_dict(make_ppo_state_dict(pc))
    alg = PPO(ac, num_learning_epochs=pc.epochs, num_mini_batches=1, clip_param=0.2, gamma=pc.gamma, lam=pc.lam,
              value_loss_coef=1.0, entropy_coef=pc.entropy_coef, learning_rate=pc.lr, max_grad_norm=1.0,
              use_clipped_value_loss=True, schedule="adaptive", desired_kl=0.01, device="cuda", **kw)
    alg.init_storage(pc.num_envs, pc.num_steps, [pc.obs_dim], [None], [pc.act_dim])
    return ac, alg


def test_state_dict_keys_match_reference():
    pc, g = PPO_CASE, load()
    ac, _ = make_alg(pc)
    ref_keys = sorted(k[len("final."):] for k in g if k.startswith("final."))
    assert sorted(ac.state_dict().keys()) == ref_keys


def test_act_and_rollout_quantities():
    pc, g, ro = PPO_CASE, load(), make_ppo_rollout(PPO_CASE)
    ac, alg = make_alg(pc)
    with torch.inference_mode():
        for s in range(pc.num_steps):
            obs = ro["obs"][s].cuda()
            a = alg.act(obs, obs)
            assert a.shape == (pc.num_envs, pc.act_dim)
            assert rel_err(ac.action_mean.cpu(), g["mu"][s]) < 1e-4
            assert rel_err(ac.action_std.cpu(), g["sigma"][s]) < 1e-6
            assert rel_err(alg.transition.values.cpu(), g["values"][s]) < 1e-5
            lp = ac.get_actions_log_prob(torch.from_numpy(g["actions"][s]).cuda())
            assert rel_err(lp.cpu(), g["actions_log_prob"][s].reshape(-1)) < 1e-4
            alg.process_env_step(ro["rewards"][s].cuda(), ro["dones"][s].cuda(), {})
    assert alg.storage.step == pc.num_steps
    with pytest.raises(AssertionError):
        alg.transition.observations = ro["obs"][0].cuda()
        alg.transition.critic_observations = ro["obs"][0].cuda()
        alg.storage.add_transitions(alg.transition)      # "Rollout buffer overflow"


def fill_storage(alg, pc, g, ro):
    st = alg.storage
    st.observations.copy_(ro["obs"])
    st.actions.copy_(torch.from_numpy(g["actions"]))
    st.rewards.copy_(ro["rewards"].unsqueeze(-1))
    st.dones.copy_(ro["dones"].unsqueeze(-1))
    st.values.copy_(torch.from_numpy(g["values"]))
    st.actions_log_prob.copy_(torch.from_numpy(g["actions_log_prob"]))
    st.mu.copy_(torch.from_numpy(g["mu"]))
    st.sigma.copy_(torch.from_numpy(g["sigma"]))
    st.step = pc.num_steps


@pytest.mark.parametrize("mode", ["fused_graph", "fused_eager", "generic"])
def test_gae_and_update_match_reference(mode):
    pc, g, ro = PPO_CASE, load(), make_ppo_rollout(PPO_CASE)
    ac, alg = make_alg(pc, fused=mode != "generic", use_cuda_graph=mode == "fused_graph")
    fill_storage(alg, pc, g, ro)
    alg.compute_returns(ro["last_obs"].cuda())
    assert rel_err(alg.storage.returns.cpu(), g["returns"]) < 1e-5
    assert rel_err(alg.storage.advantages.cpu(), g["advantages"]) < 1e-4
    vl, sl = alg.update()
    assert abs(vl / g["mean_value_loss"] - 1) < 1e-4
    assert abs(sl - g["mean_surrogate_loss"]) < 1e-4 * max(1.0, abs(g["mean_surrogate_loss"]))
    assert abs(alg.learning_rate / g["final_lr"] - 1) < 1e-6
    if mode == "fused_graph":
        assert alg.use_cuda_graph and len(alg._graphs) == 1      # the graph path really ran
    sd0 = make_ppo_state_dict(pc)
    for k, v in ac.state_dict().items():
        ref = torch.from_numpy(g["final." + k])
        if k == "std":
            assert torch.equal(v.cpu(), ref)
            continue
        d_ref = ref - sd0[k]
        d = v.cpu() - sd0[k]
        frac_bad = ((d - d_ref).abs() > 0.02 * pc.lr).float().mean().item()
        assert frac_bad < 0.005, (k, frac_bad)


def test_gae_kernel_large_matches_oracle():
    from fullysnnforleggedrobots_b200 import gae
    from oracle import snn_oracle as O
    gen = torch.Generator().manual_seed(5)
    S, N = 24, 4096
    rewards = torch.randn(S, N, generator=gen)
    values = torch.randn(S, N, generator=gen)
    dones = (torch.rand(S, N, generator=gen) < 0.05).to(torch.uint8)
    last = torch.randn(N, generator=gen)
    ret_ref, adv_ref = O.gae_returns(rewards.unsqueeze(-1), dones.unsqueeze(-1), values.unsqueeze(-1),
                                     last.unsqueeze(-1), 0.99, 0.95)
    ret, adv = gae(rewards.cuda(), dones.cuda(), values.cuda(), last.cuda(), 0.99, 0.95)
    assert rel_err(ret.cpu(), ret_ref.squeeze(-1)) < 1e-6
    assert rel_err(adv.cpu(), adv_ref.squeeze(-1)) < 1e-5


def test_second_update_replays_graph_and_stays_finite():
    pc, g, ro = PPO_CASE, load(), make_ppo_rollout(PPO_CASE)
    ac, alg = make_alg(pc)
    for _ in range(3):
        fill_storage(alg, pc, g, ro)
        alg.compute_returns(ro["last_obs"].cuda())
        vl, sl = alg.update()
        assert vl == vl and sl == sl
    assert all(torch.isfinite(p).all() for p in ac.parameters())
    sd = alg.optimizer.state_dict()
    assert len(sd["state"]) == len(alg.optimizer.params)


@pytest.mark.parametrize("dims,B", [((48, 256, 256), 300), ((20, 512, 256, 128), 131), ((7, 24), 5)])
def test_tensor_core_critic_matches_torch_fp32(dims, B):
    """snn_mlp_fwd / snn_mlp_bwd vs a plain PyTorch fp32 critic (reference: nn.Sequential Linear/ELU)."""
    from fullysnnforleggedrobots_b200 import CriticEngine
    torch.manual_seed(3)
    layers = []
    for i in range(len(dims) - 1):
        layers += [torch.nn.Linear(dims[i], dims[i + 1]), torch.nn.ELU()]
    layers.append(torch.nn.Linear(dims[-1], 1))
    seq = torch.nn.Sequential(*layers).cuda()
    ce = CriticEngine.from_sequential(seq)
    assert ce is not None
    x = torch.randn(B, dims[0], device="cuda") * 1.5
    gv = torch.randn(B, device="cuda")
    value, trace = ce.forward(x)
    ref = seq(x)
    assert rel_err(value.cpu(), ref.detach().squeeze(-1).cpu()) < 1e-4
    ref.backward(gv.view_as(ref))
    gw, gb = ce.backward(gv, trace, B)
    lin = [m for m in seq if isinstance(m, torch.nn.Linear)]
    for i, l in enumerate(lin):
        assert rel_err(gw[i].cpu(), l.weight.grad.cpu()) < 1e-4, f"dW{i}"
        assert rel_err(gb[i].cpu(), l.bias.grad.cpu()) < 1e-4, f"db{i}"


@pytest.mark.parametrize("dp_mode", ["single", "segments"])
def test_data_parallel_graph_modes_match_single_process(dp_mode, monkeypatch):
    """The data-parallel update with a world of one rank must reproduce the single-process fused update: same kernels,
    same order, one extra (identity) collective.  "single": the NCCL all-reduce captured inside the step graph;
    "segments": graph A -> eager NCCL all-reduce -> graph B."""
    monkeypatch.setenv("SNN_DP_GRAPH", dp_mode)
    import socket

    import torch.distributed as dist

    from fullysnnforleggedrobots_b200 import dist as snn_dist
    pc, g, ro = PPO_CASE, load(), make_ppo_rollout(PPO_CASE)

    def run(reducer):
        torch.manual_seed(0)
        from fullysnnforleggedrobots_b200 import PPO, ActorCriticCassie
        ac = ActorCriticCassie(pc.obs_dim, pc.obs_dim, pc.act_dim, actor_hidden_dims=list(pc.hidden),
                               critic_hidden_dims=list(pc.critic_hidden), activation="elu", init_noise_std=1.0)
        ac.load_state_dict(make_ppo_state_dict(pc))
        alg = PPO(ac, num_learning_epochs=pc.epochs, num_mini_batches=1, clip_param=0.2, gamma=pc.gamma, lam=pc.lam,
                  value_loss_coef=1.0, entropy_coef=pc.entropy_coef, learning_rate=pc.lr, max_grad_norm=1.0,
                  use_clipped_value_loss=True, schedule="adaptive", desired_kl=0.01, device="cuda", data_parallel=reducer)
        alg.init_storage(pc.num_envs, pc.num_steps, [pc.obs_dim], [None], [pc.act_dim])
        fill_storage(alg, pc, g, ro)
        alg.compute_returns(ro["last_obs"].cuda())
        out = alg.update()
        return out, {k: v.detach().cpu().clone() for k, v in ac.state_dict().items()}, alg

    (vl0, sl0), sd0, _ = run(None)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1)
    try:
        (vl1, sl1), sd1, alg = run(snn_dist.GradientAllReducer())
        assert alg.use_cuda_graph and alg.dp_graph_mode == dp_mode, alg.dp_graph_note
        if dp_mode == "single":
            assert ("full", 0) in alg._graphs
        else:
            assert ("pre", 0) in alg._graphs and "post" in alg._graphs
    finally:
        dist.destroy_process_group()
    assert abs(vl0 - vl1) <= 1e-5 * abs(vl0) and abs(sl0 - sl1) <= 1e-5 * max(1.0, abs(sl0))
    for k in sd0:
        assert rel_err(sd1[k], sd0[k]) < 1e-5, k


def test_act_cuda_graph_is_consistent_and_sees_updated_weights():
    pc, g, ro = PPO_CASE, load(), make_ppo_rollout(PPO_CASE)
    ac, alg = make_alg(pc)
    obs = ro["obs"][0].cuda()
    with torch.inference_mode():
        for k in range(6):                                    # 2 eager warm-ups, capture, replays
            a = alg.act(obs, obs)
            assert rel_err(alg.transition.action_mean.cpu(), g["mu"][0]) < 1e-4
            assert rel_err(alg.transition.values.cpu(), g["values"][0]) < 1e-5
            lp = ac.get_actions_log_prob(a)                   # distribution state of the eager module is stale: recompute
            z = (a - alg.transition.action_mean) / alg.transition.action_sigma
            lp_ref = (-0.5 * z * z - torch.log(alg.transition.action_sigma) - 0.9189385332).sum(-1)
            assert rel_err(alg.transition.actions_log_prob.cpu(), lp_ref.cpu()) < 1e-4
            alg.transition.clear()
        assert alg._act_state["graph"] is not None
        a1 = alg.act(obs, obs).clone()
        a2 = alg.act(obs, obs).clone()
        assert not torch.equal(a1, a2)                        # fresh noise on every replay
    # parameters change -> the replayed graph must use the new weights
    fill_storage(alg, pc, g, ro)
    alg.compute_returns(ro["last_obs"].cuda())
    alg.update()
    with torch.inference_mode():
        alg.act(obs, obs)
        mu_graph = alg.transition.action_mean.clone()
        mu_eager = ac.act_inference(obs)
        v_graph = alg.transition.values.clone()
        v_eager = ac.evaluate(obs)
    assert rel_err(mu_graph.cpu(), mu_eager.cpu()) < 1e-5
    assert rel_err(mu_graph.cpu(), g["mu"][0]) > 1e-6         # and they did change
    # the critic branch of the graph (tensor-core MLP kernels, own operand planes) follows update() ...
    assert alg._act_state["ce"] is not None
    assert rel_err(v_graph.cpu(), v_eager.cpu()) < 1e-4
    assert rel_err(v_graph.cpu(), g["values"][0]) > 1e-6
    # ... and a parameter written through torch (load_state_dict, manual edits)
    with torch.no_grad():
        ac.critic[0].bias.add_(0.25)
    with torch.inference_mode():
        alg.act(obs, obs)
        v2 = alg.transition.values.clone()
        v2_eager = ac.evaluate(obs)
    assert rel_err(v2.cpu(), v2_eager.cpu()) < 1e-4 and rel_err(v2.cpu(), v_graph.cpu()) > 1e-4


def test_act_torch_critic_branch_switch(monkeypatch):
    """SNN_ACT_CRITIC=torch keeps the rollout's critic on torch (cuBLAS); same values within fp32 noise."""
    monkeypatch.setenv("SNN_ACT_CRITIC", "torch")
    pc, g, ro = PPO_CASE, load(), make_ppo_rollout(PPO_CASE)
    ac, alg = make_alg(pc)
    obs = ro["obs"][0].cuda()
    with torch.inference_mode():
        for k in range(4):
            alg.act(obs, obs)
            assert rel_err(alg.transition.values.cpu(), g["values"][0]) < 1e-5
            alg.transition.clear()
    assert alg._act_state["graph"] is not None and alg._act_state["ce"] is None


def test_permute_once_gathers_every_field_like_index_select():
    """snn_gather_rows (one launch for all fields) against torch.index_select, odd row widths and a ragged size."""
    from fullysnnforleggedrobots_b200 import RolloutStorage
    S, N, O, A = 5, 37, 7, 3
    st = RolloutStorage(N, S, [O], [11], [A], device="cuda")
    g = torch.Generator(device="cuda").manual_seed(3)
    for name in ("observations", "privileged_observations", "actions", "values", "returns", "actions_log_prob",
                 "advantages", "mu", "sigma"):
        t = getattr(st, name)
        t.copy_(torch.randn(t.shape, device="cuda", generator=g))
    torch.manual_seed(11)
    f, mb = st.permute_once(4)
    torch.manual_seed(11)
    idx = torch.randperm(4 * mb, device="cuda")
    assert mb == S * N // 4
    pairs = {"obs": st.observations, "critic_obs": st.privileged_observations, "actions": st.actions, "values": st.values,
             "returns": st.returns, "logp": st.actions_log_prob, "adv": st.advantages, "mu": st.mu, "sigma": st.sigma}
    for k, src in pairs.items():
        assert torch.equal(f[k], src.flatten(0, 1).index_select(0, idx)), k


@pytest.mark.parametrize("B,A", [(4096, 12), (37, 3), (1, 1), (300, 64)])
def test_sample_logprob_kernel_matches_torch_normal(B, A):
    """snn_sample_logprob (the rollout step's sampling / log-prob / sigma tail) against torch.distributions.Normal
    (actor_critic.py:398-421) fed the same noise."""
    import ctypes as C

    from fullysnnforleggedrobots_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(B + A)
    mu = torch.randn(B, A, device="cuda", generator=g) * 2
    log_std = torch.randn(A, device="cuda", generator=g) * 0.3 - 0.2
    noise = torch.randn(B, A, device="cuda", generator=g)
    actions, logp, sigma = torch.empty(B, A, device="cuda"), torch.empty(B, device="cuda"), torch.empty(B, A, device="cuda")
    p = lambda t: C.c_void_p(t.data_ptr())
    _lib.check(_lib.lib().snn_sample_logprob(p(mu), p(log_std), p(noise), B, A, p(actions), p(logp), p(sigma),
                                             C.c_void_p(torch.cuda.current_stream().cuda_stream)), "snn_sample_logprob")
    std = torch.exp(log_std)
    dist = torch.distributions.Normal(mu, mu * 0. + std, validate_args=False)
    assert rel_err(sigma.cpu(), dist.stddev.cpu()) < 1e-6
    assert rel_err(actions.cpu(), (mu + std * noise).cpu()) < 1e-6
    assert rel_err(logp.cpu(), dist.log_prob(actions).sum(-1).cpu()) < 1e-5


def _two_updates(pc, g, ro, graph, ent2, intermezzo=False):
    ac, alg = make_alg(pc, use_cuda_graph=graph)
    out = []
    for k in range(2):
        fill_storage(alg, pc, g, ro)
        alg.compute_returns(ro["last_obs"].cuda())
        if k == 1:
            alg.entropy_coef = ent2                 # a runner annealing the entropy bonus between iterations
            alg.clip_param = 0.15
        torch.manual_seed(1000 + k)                 # the update's randperm: same row order in the runs being compared
        out.append(alg.update())
        if intermezzo and k == 0:
            # eager calls with a LARGER batch between two graph replays: they re-allocate the engines' internal
            # workspaces, which the captured graphs must not depend on (engine.py:68)
            big = torch.randn(4096, pc.obs_dim, device="cuda")
            with torch.inference_mode():
                ac.act_inference(big)
            x = big[:1024].clone().requires_grad_(True)
            m, _ = ac.actor(x)
            m.sum().backward()
            for p in ac.parameters():
                if p.grad is not None:
                    p.grad.zero_()
            torch.cuda.synchronize()
    return out, {k: v.detach().cpu().clone() for k, v in ac.state_dict().items()}, alg


def test_graphs_follow_hyperparameter_changes_between_updates():
    """ADVICE r1 (ppo.py:486): the captured minibatch-step graphs bake in clip_param / entropy_coef / ...; the
    reference reads them every step.  Changing them between two updates must give exactly what eager launches give."""
    pc, g, ro = PPO_CASE, load(), make_ppo_rollout(PPO_CASE)
    out_e, sd_e, _ = _two_updates(pc, g, ro, graph=False, ent2=0.2)
    out_g, sd_g, alg = _two_updates(pc, g, ro, graph=True, ent2=0.2)
    assert alg.use_cuda_graph and len(alg._graphs) == 1
    _, sd_same, _ = _two_updates(pc, g, ro, graph=True, ent2=pc.entropy_coef)
    for k in sd_e:
        assert torch.equal(sd_g[k], sd_e[k]), k                   # same kernels, same order: bit-identical
    assert any(not torch.equal(sd_g[k], sd_same[k]) for k in sd_e)            # and the change really mattered
    assert out_g == out_e


def test_graphs_survive_larger_eager_calls_and_new_storage():
    """ADVICE r1 (engine.py:68): graphs own their workspaces; a larger eager call in between must not disturb a replay.
    Also: init_storage() again (new minibatch buffers) invalidates the graphs instead of replaying over freed memory."""
    pc, g, ro = PPO_CASE, load(), make_ppo_rollout(PPO_CASE)
    _, sd_ref, _ = _two_updates(pc, g, ro, graph=True, ent2=pc.entropy_coef)
    _, sd_int, alg = _two_updates(pc, g, ro, graph=True, ent2=pc.entropy_coef, intermezzo=True)
    for k in sd_ref:
        assert torch.equal(sd_int[k], sd_ref[k]), k
    sig0 = alg._graph_sig
    alg.init_storage(pc.num_envs, pc.num_steps, [pc.obs_dim], [None], [pc.act_dim])
    assert alg._graphs == {} and alg._graph_sig is None
    fill_storage(alg, pc, g, ro)
    alg.compute_returns(ro["last_obs"].cuda())
    vl, sl = alg.update()
    assert vl == vl and alg._graph_sig is not None and len(alg._graphs) == 1
    assert alg._graph_sig[:-1] == sig0[:-1]                 # same hyperparameters, (possibly) other buffer addresses


def test_act_graph_static_buffers_accept_other_modes_and_strides():
    """ADVICE r1 (ppo.py:240): the rollout graph's static buffers must not be inference tensors or inherit strides."""
    pc, g, ro = PPO_CASE, load(), make_ppo_rollout(PPO_CASE)
    ac, alg = make_alg(pc)
    obs = ro["obs"][0].cuda()
    with torch.inference_mode():
        for _ in range(3):
            alg.act(obs, obs)
            alg.transition.clear()
    assert alg._act_state["graph"] is not None
    wide = torch.randn(pc.num_envs, 2 * pc.obs_dim, device="cuda")
    view = wide[:, ::2]                                     # non-contiguous observations, plain no_grad mode
    assert not view.is_contiguous()
    with torch.no_grad():
        alg.act(view, view)
        mu_graph = alg.transition.action_mean.clone()
        mu_eager = ac.act_inference(view.contiguous())
    assert alg._act_state["graph"] is not None and alg.act_cuda_graph
    assert rel_err(mu_graph.cpu(), mu_eager.cpu()) < 1e-6


@pytest.mark.parametrize("graph", [True, False])
def test_fused_steps_overwrite_every_gradient_they_use(graph):
    """The fused path clears the gradient bucket once per update(), not per minibatch step: every fused kernel must then
    OVERWRITE its slice.  Poison the bucket right after that clear (everything but the reference's unused `std`): the
    update must come out exactly as from a clean bucket."""
    pc, g, ro = PPO_CASE, load(), make_ppo_rollout(PPO_CASE)
    res = []
    for poison in (False, True):
        ac, alg = make_alg(pc, use_cuda_graph=graph)
        fill_storage(alg, pc, g, ro)
        alg.compute_returns(ro["last_obs"].cuda())
        opt = alg.optimizer
        if poison:
            clear = opt.zero_grad

            def poisoned(*a, **k):
                clear(*a, **k)
                for name, p in ac.named_parameters():
                    if name != "std":
                        p.grad.fill_(float("nan"))
                opt.flat_grads[opt.n:].fill_(float("nan"))
            opt.zero_grad = poisoned
        torch.manual_seed(7)
        res.append((alg.update(), {k: v.detach().cpu().clone() for k, v in ac.state_dict().items()}))
    assert res[0][0] == res[1][0]
    for k in res[0][1]:
        assert torch.isfinite(res[1][1][k]).all(), k
        assert torch.equal(res[0][1][k], res[1][1][k]), k


def test_update_is_bitwise_reproducible():
    """Two updates from identical states give bit-identical parameters: every reduction of the fused step has a fixed
    order (per-tile partial rows + split-K tiles of the actor, per-(row tile, lane quarter) partial rows of the critic's
    bias gradients — shared-memory float atomics until r2 —, block partials of the PPO head and the gradient norm)."""
    pc, g, ro = PPO_CASE, load(), make_ppo_rollout(PPO_CASE)
    res = []
    for _ in range(2):
        ac, alg = make_alg(pc)
        fill_storage(alg, pc, g, ro)
        alg.compute_returns(ro["last_obs"].cuda())
        torch.manual_seed(11)
        out = alg.update()
        res.append((out, {k: v.detach().cpu().clone() for k, v in ac.state_dict().items()}))
    assert res[0][0] == res[1][0]
    for k in res[0][1]:
        assert torch.equal(res[0][1][k], res[1][1][k]), k
