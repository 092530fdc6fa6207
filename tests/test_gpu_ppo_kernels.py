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
