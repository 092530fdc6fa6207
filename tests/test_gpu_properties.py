"""Size-independent properties of the CUDA path at BASELINE.json's full sizes (C1 dims: obs 48, hidden 256-256,
12 actions, T=5; B = 4096 rollout step and B = 24576 PPO minibatch), where the CPU oracle is too slow to be the
checker: row independence, run-to-run reproducibility, linearity of BPTT in the upstream gradient, the decoder as
a checksum of the recorded spike trains, sub-sampled oracle parity, and the edge cases B = 0 / B = 1."""
import math

import pytest
import torch

from helpers import rel_err

pytestmark = pytest.mark.gpu


def make_actor(seed=0, **kw):
    from fullysnnforleggedrobots_b200 import PopSpikeActor
    torch.manual_seed(seed)
    return PopSpikeActor(48, 12, 10, 10, [256, 256], (-5, 5), math.sqrt(0.15), spike_ts=5, **kw).cuda()


def params_of(actor):
    layers = actor.snn.linear_layers()
    return (actor.encoder.mean.data, actor.encoder.std.data, [l.weight.data for l in layers],
            [l.bias.data for l in layers], actor.decoder.decoder.weight.data, actor.decoder.decoder.bias.data)


def grads_of(actor, obs, G):
    for p in actor.parameters():
        p.grad = None
    mu, _ = actor(obs)
    (mu * G).sum().backward()
    return mu.detach(), {n: p.grad.detach().clone() for n, p in actor.named_parameters() if p.grad is not None}


@pytest.mark.parametrize("B", [4096, 24576])
def test_rows_are_independent_and_runs_reproducible(B):
    actor = make_actor()
    g = torch.Generator(device="cuda").manual_seed(1)
    obs = torch.randn(B, 48, device="cuda", generator=g)
    with torch.inference_mode():
        mu_full, _ = actor(obs)
        mu_again, _ = actor(obs)
        mu_part, _ = actor(obs[1000:1777].contiguous())         # ragged sub-batch, different tiling
    assert torch.equal(mu_full, mu_again)
    assert torch.equal(mu_full[1000:1777], mu_part)
    assert torch.isfinite(mu_full).all()


def test_forward_backward_is_bitwise_reproducible_at_full_size():
    """Every reduction has a fixed order (per-thread sums, split-K partials reduced in order), and every operand the
    tensor core reads is written through the async proxy: two runs on the same input must agree bit for bit.  (This is
    the test that caught stale spike-operand reads when the expanders wrote the operand slot with plain st.shared.)"""
    B = 24576
    actor = make_actor()
    g = torch.Generator(device="cuda").manual_seed(3)
    obs = torch.randn(B, 48, device="cuda", generator=g)
    G = torch.randn(B, 12, device="cuda", generator=g)
    mu0, g0 = grads_of(actor, obs, G)
    for _ in range(3):
        mu1, g1 = grads_of(actor, obs, G)
        assert torch.equal(mu0, mu1)
        for n in g0:
            assert torch.equal(g0[n], g1[n]), n


def test_training_forward_equals_inference_forward_and_trace_checksum():
    B = 24576
    actor = make_actor()
    obs = torch.randn(B, 48, device="cuda")
    with torch.inference_mode():
        mu_inf, _ = actor(obs)
    mu, trace = actor.engine.forward(obs, *params_of(actor), train=True)
    assert torch.equal(mu, mu_inf)
    tr = actor.engine.unpack_trace(trace, B)
    # decoder recomputed on the host from the recorded output spike train == mu (checksum of the trace)
    rate = tr["spikes"][-1].sum(0) / 5.0                                              # [B, 120]
    w = actor.decoder.decoder.weight.data.cpu().reshape(12, 10)
    mu_chk = (rate.reshape(B, 12, 10) * w).sum(-1) + actor.decoder.decoder.bias.data.cpu()
    assert rel_err(mu.cpu(), mu_chk) < 1e-5
    # spikes are exactly the thresholded voltages, layer by layer
    for l in range(3):
        assert torch.equal(tr["spikes"][l], (tr["volt"][l] > 0.5).float())
    assert 0.02 < tr["enc_spikes"].mean().item() < 0.15


def test_bptt_is_linear_in_upstream_gradient_and_reproducible():
    B = 24576
    actor = make_actor()
    obs = torch.randn(B, 48, device="cuda")
    G1 = torch.randn(B, 12, device="cuda")
    G2 = torch.randn(B, 12, device="cuda")
    _, g1 = grads_of(actor, obs, G1)
    _, g2 = grads_of(actor, obs, G2)
    _, g12 = grads_of(actor, obs, G1 + 2.0 * G2)
    _, g1b = grads_of(actor, obs, G1)
    for k in g1:
        assert rel_err(g12[k], g1[k] + 2.0 * g2[k]) < 2e-4, k
        if k.endswith("weight") or "encoder" in k:
            assert torch.equal(g1[k], g1b[k]), f"{k} not bit-reproducible"      # fixed-order split-K reduction
        else:
            assert rel_err(g1[k], g1b[k]) < 1e-5, k


def _tie_rows(tr_gpu, tr_ref):
    """Rows (samples) whose spike trains differ anywhere between the kernels and the oracle: a membrane value within fp32
    rounding of a threshold is decided either way by two correct implementations.  [rows] bool."""
    bad = (tr_gpu["enc_spikes"] != tr_ref.enc_spikes).any(dim=2).any(dim=0)
    for a, b in zip(tr_gpu["spikes"], tr_ref.spikes):
        bad |= (a != b).any(dim=2).any(dim=0)
    return bad


@pytest.mark.parametrize("B", [24576, 4096, 1000])
def test_full_size_subsample_matches_oracle(B):
    """Oracle parity on a 192-row sub-sample of a PPO minibatch (24576: 48-sample tiles) and of rollout-sized batches
    (4096 / 1000: the plan picks narrower 32- / 16-sample tiles there).  Rows with a threshold tie (spike trains differ
    from the oracle's) are identified from the trace and taken out of the loss on both sides; every other row's forward
    and the gradients of the loss over those rows are compared unconditionally."""
    from oracle import snn_oracle as O
    actor = make_actor()
    eng = actor.engine
    obs = torch.randn(B, 48, device="cuda")
    idx = torch.arange(B // 5, B // 5 + 192)
    sd = {"actor." + k: v.detach().cpu() for k, v in actor.state_dict().items()}
    p = O.ActorParams.from_state_dict(sd, spike_ts=5)
    obs_s = obs[idx.cuda()].cpu()
    mu_ref, _, tr = O.actor_forward(obs_s, p, want_traces=True)
    mu, trace = eng.forward(obs, *params_of(actor), train=True)
    ties = _tie_rows(eng.unpack_trace(trace, B, rows=idx), tr)
    assert float(ties.double().mean()) <= 0.01
    assert rel_err(mu[idx.cuda()].cpu()[~ties], mu_ref[~ties]) < 1e-4
    Gs = torch.randn(192, 12)
    Gs[ties] = 0.0
    G = torch.zeros(B, 12, device="cuda")
    G[idx.cuda()] = Gs.cuda()
    grads = eng.backward(obs, mu, G, trace, *params_of(actor), need_dobs=False)
    gr = O.actor_backward(obs_s, p, tr, Gs, need_dobs=False)
    for l in range(3):
        assert rel_err(grads["weights"][l].cpu(), gr["weights"][l]) < 1e-4, f"dW{l}"
        assert rel_err(grads["biases"][l].cpu(), gr["biases"][l]) < 1e-4, f"db{l}"
    assert rel_err(grads["enc_mean"].cpu(), gr["enc_mean"]) < 1e-4
    assert rel_err(grads["enc_std"].cpu(), gr["enc_std"]) < 1e-4
    assert rel_err(grads["dec_w"].cpu(), gr["dec_w"]) < 1e-4


def test_edge_batches():
    actor = make_actor()
    with torch.inference_mode():
        mu0, std = actor(torch.empty(0, 48, device="cuda"))
        mu1, _ = actor(torch.zeros(1, 48, device="cuda"))
    assert mu0.shape == (0, 12) and std.shape == (12,)
    assert mu1.shape == (1, 12) and torch.isfinite(mu1).all()
    with pytest.raises(RuntimeError):
        actor(torch.zeros(4, 47, device="cuda"))          # wrong observation width


@pytest.mark.parametrize("planes", [1, 2])
def test_reduced_plane_modes_stay_close(planes):
    """fwd_planes < 3 (approximate weights) is a speed option, not the parity mode: spikes may flip, but rarely."""
    B = 4096
    ref = make_actor(seed=5)
    fast = make_actor(seed=5, fwd_planes=planes)
    obs = torch.randn(B, 48, device="cuda")
    _, t_ref = ref.engine.forward(obs, *params_of(ref), train=True)
    _, t_fast = fast.engine.forward(obs, *params_of(fast), train=True)
    a, b = ref.engine.unpack_trace(t_ref, B), fast.engine.unpack_trace(t_fast, B)
    mm = max((a["spikes"][l] != b["spikes"][l]).float().mean().item() for l in range(3))
    assert mm < (2e-2 if planes == 1 else 1e-4)


@pytest.mark.parametrize("T,B", [(10, 300), (20, 257), (32, 64), (1, 40), (3, 170), (4, 200), (8, 130), (16, 100)])
def test_other_spike_timestep_counts_match_oracle(T, B):
    """BASELINE.json configs[4] sweeps T = 5 / 10 / 20; the kernels support 1..32.  The batch-tile geometry depends on T
    (csrc/plan.cuh): T = 4 / 8 / 16 fill 256 accumulator columns exactly (four full 64-column operand blocks),
    T > 16 uses a single accumulator set and two MMA sub-tiles, several (T, B) leave a ragged last tile."""
    from fullysnnforleggedrobots_b200 import PopSpikeActor
    from oracle import snn_oracle as O
    torch.manual_seed(T)
    actor = PopSpikeActor(12, 3, 10, 10, [64, 48], (-5, 5), math.sqrt(0.15), spike_ts=T).cuda()
    eng = actor.engine
    obs = torch.randn(B, 12)
    sd = {"actor." + k: v.detach().cpu() for k, v in actor.state_dict().items()}
    p = O.ActorParams.from_state_dict(sd, spike_ts=T)
    mu_ref, _, tr = O.actor_forward(obs, p, want_traces=True)
    mu, trace = eng.forward(obs.cuda(), *params_of(actor), train=True)
    ties = _tie_rows(eng.unpack_trace(trace, B), tr)      # rows with a threshold tie leave the loss on both sides
    assert float(ties.double().mean()) <= 0.01
    assert rel_err(mu.cpu()[~ties], mu_ref[~ties]) < 1e-4
    G = torch.randn(B, 3)
    G[ties] = 0.0
    grads = eng.backward(obs.cuda(), mu, G.cuda(), trace, *params_of(actor), need_dobs=True)
    gr = O.actor_backward(obs, p, tr, G, need_dobs=True)
    for l in range(3):
        assert rel_err(grads["weights"][l].cpu(), gr["weights"][l]) < 1e-4, f"dW{l}"
        assert rel_err(grads["biases"][l].cpu(), gr["biases"][l]) < 1e-4, f"db{l}"
    assert rel_err(grads["enc_mean"].cpu(), gr["enc_mean"]) < 1e-4
    assert rel_err(grads["obs"].cpu()[~ties], gr["obs"][~ties]) < 1e-4
    if ties.any():
        assert grads["obs"].cpu()[ties].abs().max().item() == 0.0      # rows are independent: no loss, no gradient
