"""Direct tests of the PPO-head / optimizer kernels at the sizes bench.py runs them: in round 1
snn_ppo_head, snn_clip_adam (sumsq + clip + Adam) and snn_lr_adapt were only exercised through a 32-sample, one-CTA
PPO fixture.  Here: B = 24 576 (96 blocks + finalize), A = 12 / 16, n = 303 k parameters, both clip branches, against
plain PyTorch fp32 (autograd for the head's closed-form gradients, torch.optim.Adam + clip_grad_norm_ for the step).
Reference: AMP-rl/algorithms/ppo.py:132-177."""
import ctypes as C

import pytest
import torch

from helpers import rel_err

pytestmark = pytest.mark.gpu


def _p(t):
    return C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def torch_head(mu, log_std, value, actions, old_logp, adv, returns, target_v, old_mu, old_sigma, clip, vcoef, ecoef,
               clipped_value):
    """The reference's loss block (ppo.py:132-171) on leaves mu / log_std / value -> (grads, stats)."""
    mu = mu.clone().requires_grad_(True)
    log_std = log_std.clone().requires_grad_(True)
    value = value.clone().requires_grad_(True)
    sigma = torch.exp(log_std)
    dist = torch.distributions.Normal(mu, mu * 0. + sigma, validate_args=False)
    logp = dist.log_prob(actions).sum(-1)
    entropy = dist.entropy().sum(-1)
    with torch.no_grad():
        sg = dist.stddev
        kl = torch.sum(torch.log(sg / old_sigma + 1.e-5) + (old_sigma.square() + (old_mu - mu).square())
                       / (2.0 * sg.square()) - 0.5, dim=-1).mean()
    ratio = torch.exp(logp - old_logp)
    sur = torch.max(-adv * ratio, -adv * torch.clamp(ratio, 1.0 - clip, 1.0 + clip)).mean()
    if clipped_value:
        vc = target_v + (value - target_v).clamp(-clip, clip)
        vl = torch.max((value - returns).pow(2), (vc - returns).pow(2)).mean()
    else:
        vl = (returns - value).pow(2).mean()
    loss = sur + vcoef * vl - ecoef * entropy.mean()
    loss.backward()
    return mu.grad, value.grad, log_std.grad, sur.item(), vl.item(), kl.item(), entropy.mean().item()


@pytest.mark.parametrize("B,A,clipped_value", [(24576, 12, 1), (24576, 16, 0), (3072, 10, 1), (257, 3, 1)])
def test_ppo_head_matches_torch_autograd(B, A, clipped_value):
    from fullysnnforleggedrobots_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(B + A)
    r = lambda *s: torch.randn(*s, device="cuda", generator=g)
    mu, log_std = r(B, A) * 0.5, r(A) * 0.2 - 0.1
    old_mu, old_sigma = mu + r(B, A) * 0.1, torch.exp(log_std + 0.05 * r(A)).expand(B, A).contiguous()
    actions = old_mu + old_sigma * r(B, A)
    old_logp = torch.distributions.Normal(old_mu, old_sigma).log_prob(actions).sum(-1)
    adv, returns, value = r(B), r(B), r(B)
    target_v = value + 0.3 * r(B)                       # |value - target| straddles the 0.2 clip -> both branches
    clip, vcoef, ecoef = 0.2, 1.0, 0.01
    g_mu, g_value = torch.empty(B, A, device="cuda"), torch.empty(B, device="cuda")
    g_log_std, stats = torch.empty(A, device="cuda"), torch.zeros(8, device="cuda")
    scratch = torch.empty(((B + 255) // 256) * (4 + A) * 4 + 1024, dtype=torch.uint8, device="cuda")
    _lib.check(_lib.lib().snn_ppo_head(_p(mu), _p(log_std), _p(value), _p(actions), _p(old_logp), _p(adv), _p(returns),
                                       _p(target_v), _p(old_mu), _p(old_sigma), B, A, clip, vcoef, ecoef, clipped_value,
                                       _p(g_mu), _p(g_value), _p(g_log_std), _p(stats), _p(scratch), scratch.numel(),
                                       _stream()), "snn_ppo_head")
    gm, gv, gl, sur, vl, kl, ent = torch_head(mu, log_std, value, actions, old_logp, adv, returns, target_v, old_mu,
                                              old_sigma, clip, vcoef, ecoef, bool(clipped_value))
    ratio_clipped = ((torch.exp(torch.distributions.Normal(mu, torch.exp(log_std)).log_prob(actions).sum(-1) - old_logp)
                      - 1).abs() > clip).float().mean().item()
    assert 0.05 < ratio_clipped < 0.95                  # the surrogate's clipped and unclipped branches both occur
    assert rel_err(g_mu.cpu(), gm.cpu()) < 1e-4
    assert rel_err(g_value.cpu(), gv.cpu()) < 1e-5
    assert rel_err(g_log_std.cpu(), gl.cpu()) < 1e-4
    s = stats.cpu()
    assert abs(s[0].item() - sur) < 1e-5 * max(1.0, abs(sur))
    assert abs(s[1].item() - vl) < 1e-5 * max(1.0, abs(vl))
    assert abs(s[2].item() - kl) < 1e-4 * max(1.0, abs(kl))
    assert abs(s[3].item() - ent) < 1e-5 * max(1.0, abs(ent))
    assert s[4].item() == s[0].item() and s[5].item() == s[1].item() and s[6].item() == 1.0     # accumulators


@pytest.mark.parametrize("n,max_norm,wd,scale", [(303000, 1.0, 0.0, 1.0), (303000, 1e4, 1e-5, 0.125), (1237, 0.0, 0.0, 1.0),
                                                 (2_100_000, 1.0, 0.0, 0.5)])
def test_clip_adam_matches_torch(n, max_norm, wd, scale):
    """Three optimizer steps of sumsq + clip + Adam over one flat buffer vs clip_grad_norm_ + torch.optim.Adam.
    303 k = C1 actor-critic, 2.1 M = RMA actual; max_norm 1.0 clips (|g| ~ sqrt(n)), 1e4 does not, 0 disables;
    grad_scale = 1 / world is applied before the norm, as after an all-reduce(sum)."""
    from fullysnnforleggedrobots_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(n)
    prm = torch.randn(n, device="cuda", generator=g)
    ref = torch.nn.Parameter(prm.clone())
    opt = torch.optim.Adam([ref], lr=3e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=wd)
    m, v = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    lr = torch.full((1,), 3e-4, device="cuda")
    step = torch.zeros(1, dtype=torch.int32, device="cuda")
    norm = torch.zeros(1, device="cuda")
    scratch = torch.empty(4096, dtype=torch.uint8, device="cuda")
    for it in range(3):
        grads = torch.randn(n, device="cuda", generator=g) * (0.5 + it)
        _lib.check(_lib.lib().snn_clip_adam(_p(prm), _p(grads), _p(m), _p(v), n, _p(lr), _p(step), 0.9, 0.999, 1e-8, wd,
                                            max_norm, scale, _p(norm), _p(scratch), _stream()), "snn_clip_adam")
        ref.grad = grads * scale
        if max_norm > 0:
            total = torch.nn.utils.clip_grad_norm_([ref], max_norm)
            assert rel_err(norm.cpu(), total.detach().cpu().reshape(1)) < 1e-5
        opt.step()
        assert step.item() == it + 1
        # one Adam step moves an entry by ~lr: compare the MOVE, not the parameter
        assert (prm - ref.detach()).abs().max().item() < 1e-3 * 3e-4, f"step {it}"
    assert rel_err(m.cpu(), opt.state[ref]["exp_avg"].cpu()) < 1e-5
    assert rel_err(v.cpu(), opt.state[ref]["exp_avg_sq"].cpu()) < 1e-5


@pytest.mark.parametrize("kl,lr0,expect", [(0.05, 1e-3, 1e-3 / 1.5), (0.001, 1e-3, 1.5e-3), (0.01, 1e-3, 1e-3),
                                           (0.05, 1.2e-5, 1e-5), (0.001, 9e-3, 1e-2), (0.0, 1e-3, 1e-3),
                                           (-0.01, 1e-3, 1e-3)])
def test_lr_adapt_all_branches(kl, lr0, expect):
    """ppo.py:145-148: kl > 2 kl* -> lr / 1.5 (floor 1e-5); 0 < kl < kl*/2 -> lr * 1.5 (cap 1e-2); else unchanged.
    kl arrives as a SUM over `world` ranks (tail of the gradient bucket) and is scaled by 1 / world."""
    from fullysnnforleggedrobots_b200 import _lib
    for world in (1, 8):
        k = torch.full((1,), kl * world, device="cuda")
        lr = torch.full((1,), lr0, device="cuda")
        _lib.check(_lib.lib().snn_lr_adapt(_p(k), 1.0 / world, 0.01, _p(lr), _stream()), "snn_lr_adapt")
        assert abs(lr.item() / expect - 1) < 1e-6


@pytest.mark.parametrize("B,dims,tanh", [(24576, (48, 256, 256, 12), 0), (2050, (169, 512, 256, 128, 10), 1)])
def test_split_head_matches_torch_head_and_one_shot_backward(B, dims, tanh):
    """The fused PPO step's split head (snn_bwd_ppo_top + snn_ppo_value_loss + snn_ppo_finalize + snn_bwd_rest):
    mu bit-equal to the layered decoder, d loss / d mu / value / log_std and the loss statistics against torch autograd
    on that mu, and every actor gradient against the one-shot snn_bwd fed with the SAME d loss / d mu."""
    import math
    from fullysnnforleggedrobots_b200 import PopSpikeActor, _lib
    torch.manual_seed(B)
    O, A = dims[0], dims[-1]
    actor = PopSpikeActor(O, A, 10, 10, list(dims[1:-1]), (-5, 5), math.sqrt(0.15), spike_ts=5, dec_tanh=bool(tanh)).cuda()
    eng = actor.engine
    layers = actor.snn.linear_layers()
    params = (actor.encoder.mean.data, actor.encoder.std.data, [l.weight.data for l in layers], [l.bias.data for l in layers],
              actor.decoder.decoder.weight.data, actor.decoder.decoder.bias.data)
    g = torch.Generator(device="cuda").manual_seed(B + A)
    r = lambda *s: torch.randn(*s, device="cuda", generator=g)
    obs = r(B, O)
    mu_ref, trace = eng.forward(obs, *params, train=True)
    _, trace2 = eng.forward(obs, *params, train=True, out_mu=False, assume_packed=True)
    log_std = r(A) * 0.2 - 0.1
    old_mu, old_sigma = mu_ref + r(B, A) * 0.1, torch.exp(log_std + 0.05 * r(A)).expand(B, A).contiguous()
    actions = old_mu + old_sigma * r(B, A)
    old_logp = torch.distributions.Normal(old_mu, old_sigma).log_prob(actions).sum(-1)
    adv, returns, value = r(B), r(B), r(B)
    target_v = value + 0.3 * r(B)
    clip, vcoef, ecoef = 0.2, 1.0, 0.01
    L = _lib.lib()
    part = torch.empty(eng.ppo_head_part_bytes(B), dtype=torch.uint8, device="cuda")
    vpart = torch.empty(((B + 255) // 256) * 4, dtype=torch.uint8, device="cuda")
    ws = torch.empty(eng.workspace_bytes(B, True), dtype=torch.uint8, device="cuda")
    mu, g_mu = torch.empty(B, A, device="cuda"), torch.empty(B, A, device="cuda")
    g_value, g_log_std, stats, klo = torch.empty(B, device="cuda"), torch.empty(A, device="cuda"), torch.zeros(8, device="cuda"), torch.zeros(1, device="cuda")
    eng.backward_ppo_top(obs, trace2, params, log_std, actions, old_logp, adv, old_mu, old_sigma, clip, part, ws,
                         out_mu=mu, out_g_mu=g_mu)
    _lib.check(L.snn_ppo_value_loss(_p(value), _p(returns), _p(target_v), B, clip, vcoef, 1, _p(g_value), _p(vpart),
                                    vpart.numel(), _stream()), "snn_ppo_value_loss")
    _lib.check(L.snn_ppo_finalize(C.byref(eng.desc), B, _p(part), _p(vpart), _p(log_std), ecoef, _p(g_log_std), _p(stats),
                                  _p(klo), _stream()), "snn_ppo_finalize")
    out = {"enc_mean": torch.empty_like(params[0]), "enc_std": torch.empty_like(params[1]),
           "weights": [torch.empty_like(w) for w in params[2]], "biases": [torch.empty_like(b) for b in params[3]],
           "dec_w": torch.empty_like(params[4]), "dec_b": torch.empty_like(params[5]), "obs": None}
    eng.backward_rest(obs, trace2, params, out, ws)
    assert torch.equal(mu, mu_ref)
    gm, gv, gl, sur, vl, kl, ent = torch_head(mu_ref, log_std, value, actions, old_logp, adv, returns, target_v, old_mu,
                                              old_sigma, clip, vcoef, ecoef, True)
    assert rel_err(g_mu.cpu(), gm.cpu()) < 1e-4
    assert rel_err(g_value.cpu(), gv.cpu()) < 1e-5
    assert rel_err(g_log_std.cpu(), gl.cpu()) < 1e-4
    s = stats.cpu()
    assert abs(s[0].item() - sur) < 1e-5 * max(1.0, abs(sur))
    assert abs(s[1].item() - vl) < 1e-5 * max(1.0, abs(vl))
    assert abs(s[2].item() - kl) < 1e-4 * max(1.0, abs(kl)) and abs(klo.item() - s[2].item()) == 0.0
    assert abs(s[3].item() - ent) < 1e-5 * max(1.0, abs(ent))
    ref = eng.backward(obs, mu_ref, g_mu, trace, *params, need_dobs=False)
    for k in ("enc_mean", "enc_std", "dec_w", "dec_b"):
        assert rel_err(out[k].cpu(), ref[k].cpu()) < 2e-6, k
    for i in range(len(layers)):
        assert rel_err(out["weights"][i].cpu(), ref["weights"][i].cpu()) < 2e-6, f"dW{i}"
        assert rel_err(out["biases"][i].cpu(), ref["biases"][i].cpu()) < 2e-6, f"db{i}"


def test_clip_adam_kl_applies_the_schedule_first():
    """snn_clip_adam_kl == snn_lr_adapt followed by snn_clip_adam (same lr, same parameters, bit for bit)."""
    from fullysnnforleggedrobots_b200 import _lib
    L = _lib.lib()
    n = 303_117
    g = torch.Generator(device="cuda").manual_seed(5)
    prm0, grad = torch.randn(n, device="cuda", generator=g), torch.randn(n, device="cuda", generator=g) * 0.01
    for klv in (0.05, 0.001, 0.01):
        res = []
        for fused in (False, True):
            prm, m, v = prm0.clone(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
            lr, step = torch.full((1,), 1e-3, device="cuda"), torch.zeros(1, dtype=torch.int32, device="cuda")
            kl, scratch = torch.full((1,), klv, device="cuda"), torch.empty(4096, dtype=torch.uint8, device="cuda")
            for _ in range(3):
                if fused:
                    _lib.check(L.snn_clip_adam_kl(_p(prm), _p(grad), _p(m), _p(v), n, _p(lr), _p(step), 0.9, 0.999, 1e-8, 0.0,
                                                  1.0, 1.0, _p(kl), 1.0, 0.01, None, _p(scratch), _stream()), "snn_clip_adam_kl")
                else:
                    _lib.check(L.snn_lr_adapt(_p(kl), 1.0, 0.01, _p(lr), _stream()), "snn_lr_adapt")
                    _lib.check(L.snn_clip_adam(_p(prm), _p(grad), _p(m), _p(v), n, _p(lr), _p(step), 0.9, 0.999, 1e-8, 0.0,
                                               1.0, 1.0, None, _p(scratch), _stream()), "snn_clip_adam")
            res.append((prm, lr.clone(), step.clone()))
        assert torch.equal(res[0][1], res[1][1]) and torch.equal(res[0][2], res[1][2])
        assert torch.equal(res[0][0], res[1][0])


def test_peer_memory_allreduce_adam_matches_nccl_on_all_visible_gpus():
    """snn_clip_adam_kl_p2p (the gradient all-reduce inside the optimizer kernel, NVLink peer memory) against NCCL
    all_reduce + snn_clip_adam_kl, one process per visible GPU (tools/dp_p2p_unit.py under torchrun).  Needs >= 2 GPUs."""
    import os
    import subprocess
    import sys
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    n = min(n, 8)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr",
                        "127.0.0.1", "--master-port", "29531", os.path.join(root, "tools", "dp_p2p_unit.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "OK" in r.stdout.splitlines()[-1] and "MISMATCH" not in r.stdout, r.stdout[-2000:]
