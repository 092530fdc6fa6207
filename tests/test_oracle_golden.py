"""Pin the CPU oracle (oracle/snn_oracle.py) against fixtures produced by the unmodified reference
modules (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from cases import CASES, make_gmu, make_obs
from helpers import load_golden, mismatch_rate, oracle_params, rel_err, unpack_bits
from oracle import snn_oracle as O

FULL = [c for c in CASES if c.full]
BIG = [c for c in CASES if not c.full]


@pytest.mark.parametrize("c", FULL, ids=lambda c: c.name)
def test_forward_traces_exact(c):
    g = load_golden(c.name)
    p = oracle_params(c)
    mu, std, tr = O.actor_forward(make_obs(c), p, want_traces=True)
    # same ops in the same order on the same backend -> bit-exact
    assert torch.equal(tr.enc_spikes, torch.from_numpy(g["enc_spikes"]).float())
    for l in range(len(p.weights)):
        assert torch.equal(tr.spikes[l], torch.from_numpy(g[f"spk{l}"]).float()), f"layer {l}"
        assert rel_err(tr.volt[l], g[f"volt{l}"]) < 1e-6
    assert rel_err(mu, g["mu"]) < 1e-6
    assert rel_err(std, g["std"]) < 1e-7


@pytest.mark.parametrize("c", FULL, ids=lambda c: c.name)
def test_manual_bptt_matches_reference_autograd(c):
    g = load_golden(c.name)
    p = oracle_params(c)
    obs = make_obs(c)
    mu, std, tr = O.actor_forward(obs, p, want_traces=True)
    gr = O.actor_backward(obs, p, tr, make_gmu(c), need_dobs=True)
    L = len(p.weights)
    tol = 2e-5
    for l in range(L):
        name = f"snn.hidden_layers.{l}" if l < L - 1 else "snn.out_pop_layer"
        assert rel_err(gr["weights"][l], g[f"grad.{name}.weight"]) < tol, name
        assert rel_err(gr["biases"][l], g[f"grad.{name}.bias"]) < tol, name
    assert rel_err(gr["dec_w"], g["grad.decoder.decoder.weight"]) < tol
    assert rel_err(gr["dec_b"], g["grad.decoder.decoder.bias"]) < tol
    assert rel_err(gr["enc_mean"], g["grad.encoder.mean"]) < tol
    assert rel_err(gr["enc_std"], g["grad.encoder.std"]) < tol
    assert rel_err(gr["obs"], g["grad.obs"]) < tol


@pytest.mark.parametrize("c", FULL[:4], ids=lambda c: c.name)
def test_autograd_twin_matches_manual(c):
    p = oracle_params(c)
    obs = make_obs(c).requires_grad_(True)
    for t in [p.enc_mean, p.enc_std, p.dec_w, p.dec_b, *p.weights, *p.biases]:
        t.requires_grad_(True)
    mu, _ = O.actor_forward_autograd(obs, p)
    (mu * make_gmu(c)).sum().backward()
    g = load_golden(c.name)
    assert rel_err(mu.detach(), g["mu"]) < 1e-6
    assert rel_err(p.weights[0].grad, g["grad.snn.hidden_layers.0.weight"]) < 2e-5
    assert rel_err(obs.grad, g["grad.obs"]) < 2e-5


@pytest.mark.parametrize("c", BIG, ids=lambda c: c.name)
def test_big_dims_summaries(c):
    g = load_golden(c.name)
    p = oracle_params(c)
    obs = make_obs(c)
    mu, std, tr = O.actor_forward(obs, p, want_traces=True)
    T, B = c.spike_ts, c.batch
    assert mismatch_rate(tr.enc_spikes, unpack_bits(g["enc_spikes_bits"], tr.enc_spikes.shape)) == 0.0
    for l in range(len(p.weights)):
        assert mismatch_rate(tr.spikes[l], unpack_bits(g[f"spk{l}_bits"], tr.spikes[l].shape)) == 0.0
    assert rel_err(mu, g["mu"]) < 1e-6
    gr = O.actor_backward(obs, p, tr, make_gmu(c), need_dobs=True)
    L = len(p.weights)
    for l in range(L):
        name = f"snn.hidden_layers.{l}" if l < L - 1 else "snn.out_pop_layer"
        flat = gr["weights"][l].reshape(-1)
        samp = flat[:: max(1, flat.numel() // 4096)]
        assert rel_err(samp, g[f"gradsamp.{name}.weight"]) < 5e-5, name
        assert abs(float(flat.double().norm()) / float(g[f"gradnorm.{name}.weight"]) - 1) < 1e-5
    assert rel_err(gr["obs"], g["grad.obs"]) < 5e-5


def test_gae_matches_reference_storage():
    g = dict(np.load(__import__("os").path.join(__import__("helpers").GOLD, "gae_cassie.npz")))
    gen = torch.Generator().manual_seed(77)
    S, N = 24, 37
    rewards = torch.randn(S, N, 1, generator=gen)
    values = torch.randn(S, N, 1, generator=gen)
    dones = (torch.rand(S, N, 1, generator=gen) < 0.1).to(torch.uint8)
    last = torch.randn(N, 1, generator=gen)
    ret, adv = O.gae_returns(rewards, dones, values, last, 0.99, 0.95)
    assert rel_err(ret, g["returns"]) < 1e-6
    assert rel_err(adv, g["advantages"]) < 1e-6


@pytest.mark.skipif(not __import__("os").path.isdir("/root/reference"), reason="the reference tree is not on this box")
def test_committed_fixtures_are_what_the_reference_produces_here():
    """Re-run the unmodified reference modules (make_golden.py --check): every committed fixture must come back."""
    import os
    import subprocess
    import sys
    gold = __import__("helpers").GOLD
    r = subprocess.run([sys.executable, os.path.join(gold, "make_golden.py"), "--check"], capture_output=True,
                       text=True, timeout=600, cwd=gold)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == len(CASES) + 3        # + gae_cassie, ppo_cassie_tiny, rma_ppo_tiny
    # same torch build as the one that wrote them -> bit-identical; another build may differ in the last ulp
    assert all(("bit-identical" in l) or ("max rel diff" in l and float(l.split("max rel diff")[1].split()[0]) < 1e-5)
               for l in lines), r.stdout
