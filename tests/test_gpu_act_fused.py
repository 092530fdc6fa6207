"""The one-launch rollout forward (snn_fwd_act, csrc/act_fused.cuh) against the layered path and the oracle.

PopSpikeActor.forward under inference_mode (AMP-rl/modules/actor_critic.py:300-320) as called by PPO.act
(AMP-rl/algorithms/ppo.py:90-102).  The fused kernel and the layered path (encode -> one scan-GEMM per layer -> decode)
run the same arithmetic in the same order per neuron, but with different batch-tile widths and accumulator tilings of
the tensor core, so:
  * spikes are decided on fp32 sums whose MMA summation order is the same (K order) -> mu must be BIT-IDENTICAL;
  * the sampling tail must reproduce snn_sample_logprob bit for bit on the same noise;
  * mu must agree with the CPU oracle / the committed reference fixtures like the layered path does.
"""
import math

import pytest
import torch

from cases import CASES, make_obs, make_state_dict
from helpers import load_golden, oracle_params
from oracle import snn_oracle as O

pytestmark = pytest.mark.gpu

REL_TOL = 1e-4


def build(obs_dim, hidden, act_dim, T, variant="amp", seed=0, scale=1.5):
    from fullysnnforleggedrobots_b200 import PopSpikeActor
    torch.manual_seed(seed)
    actor = PopSpikeActor(obs_dim, act_dim, 10, 10, list(hidden), (-5, 5), math.sqrt(0.15), spike_ts=T,
                          enc_eps=1e-5 if variant == "humanoid" else 0.0, dec_tanh=variant == "cassie").cuda()
    with torch.no_grad():
        for l in actor.snn.linear_layers():
            l.weight.mul_(scale)
            l.bias.add_(0.02)
    return actor


def params_of(actor):
    layers = actor.snn.linear_layers()
    return (actor.encoder.mean.data, actor.encoder.std.data, [l.weight.data for l in layers],
            [l.bias.data for l in layers], actor.decoder.decoder.weight.data, actor.decoder.decoder.bias.data)


def both_paths(actor, obs):
    eng = actor.engine
    assert eng.act_supported(obs.shape[0]), "the fused kernel must take this geometry"
    mu_f = torch.empty(obs.shape[0], eng.desc.act_dim, device=obs.device)     # the fused kernel whatever the batch size
    eng.act(obs, *params_of(actor), None, None, mu_f, None, None, None)
    eng.fused_infer = False
    mu_l, _ = eng.forward(obs, *params_of(actor), train=False)
    eng.fused_infer = True
    torch.cuda.synchronize()
    return mu_f, mu_l


@pytest.mark.parametrize("dims,T,variant,B", [
    ((48, (256, 256), 12), 5, "amp", 1),
    ((48, (256, 256), 12), 5, "amp", 16),
    ((48, (256, 256), 12), 5, "amp", 131),
    ((48, (256, 256), 12), 5, "amp", 4096),          # BASELINE configs[0]: one 32-sample tile per CTA
    ((48, (256, 256), 12), 5, "amp", 24576),         # several tiles per CTA (persistent loop, barrier phases)
    ((38, (256, 256, 256), 10), 5, "humanoid", 512),
    ((42, (512, 256, 128), 12), 5, "amp", 1000),
    ((187, (512, 256, 128), 12), 8, "cassie", 777),  # configs[3]: T = 8, tanh decoder, K0 = 1870
    ((338, (512, 256, 128), 16), 5, "rma", 300),     # RMA-actual: K0 = 3380, two output neuron tiles
    ((9, (40, 16), 3), 16, "amp", 50),               # T = 16: the widest time axis the fused kernel takes
    ((9, (40, 16), 3), 3, "amp", 50),
])
def test_fused_mu_is_bit_identical_to_the_layered_path(dims, T, variant, B):
    actor = build(dims[0], dims[1], dims[2], T, variant, seed=B % 7)
    g = torch.Generator().manual_seed(B)
    obs = (torch.randn(B, dims[0], generator=g) * 1.5).cuda()
    mu_f, mu_l = both_paths(actor, obs)
    assert torch.isfinite(mu_f).all()
    assert mu_l.abs().max() > 0
    assert torch.equal(mu_f, mu_l), f"{(mu_f != mu_l).any(dim=1).sum().item()} of {B} rows differ"


def test_two_runs_are_bitwise_equal_at_rollout_size():
    actor = build(48, (256, 256), 12, 5, "amp", seed=3)
    obs = torch.randn(4096, 48, generator=torch.Generator().manual_seed(5)).cuda()
    eng = actor.engine
    a, _ = eng.forward(obs, *params_of(actor), train=False)
    a = a.clone()
    for _ in range(3):
        b, _ = eng.forward(obs, *params_of(actor), train=False)
        assert torch.equal(a, b)


@pytest.mark.parametrize("c", CASES, ids=lambda c: c.name)
def test_fused_forward_matches_oracle_and_reference_fixture(c):
    from test_gpu_actor_parity import build_actor
    actor = build_actor(c)
    assert actor.engine.act_supported(c.batch)
    p = oracle_params(c)
    obs = make_obs(c)
    mu_ref, _ = O.actor_forward(obs, p)
    with torch.inference_mode():
        mu, _ = actor(obs.cuda())
    bad = ((mu.cpu() - mu_ref).abs() > REL_TOL * mu_ref.abs().max()).any(dim=1).float().mean().item()
    assert bad == 0, f"{bad:.4f} of samples differ in mu"
    gold = torch.from_numpy(load_golden(c.name)["mu"])
    assert ((mu.cpu() - gold).abs() <= REL_TOL * gold.abs().max()).all()


@pytest.mark.parametrize("B,A", [(4096, 12), (37, 16), (1, 3)])
def test_sampling_tail_matches_sample_logprob_kernel_and_torch(B, A):
    import ctypes as C
    from fullysnnforleggedrobots_b200 import _lib
    actor = build(48, (256, 256), A, 5, "amp", seed=A)
    with torch.no_grad():
        actor.log_std.copy_(torch.linspace(-0.7, 0.2, A))
    obs = torch.randn(B, 48, generator=torch.Generator().manual_seed(B)).cuda()
    noise = torch.randn(B, A, generator=torch.Generator().manual_seed(B + 1)).cuda()
    eng = actor.engine
    mu = torch.empty(B, A, device="cuda")
    act, sig = torch.empty_like(mu), torch.empty_like(mu)
    logp = torch.empty(B, device="cuda")
    eng.act(obs, *params_of(actor), noise, actor.log_std.data, mu, act, logp, sig)
    mu_f, _ = eng.forward(obs, *params_of(actor), train=False)
    assert torch.equal(mu, mu_f)
    act2, sig2, logp2 = torch.empty_like(mu), torch.empty_like(mu), torch.empty(B, device="cuda")
    p = lambda t: C.c_void_p(t.data_ptr())
    _lib.check(_lib.lib().snn_sample_logprob(p(mu), p(actor.log_std.data), p(noise), B, A, p(act2), p(logp2), p(sig2),
                                             C.c_void_p(torch.cuda.current_stream().cuda_stream)), "snn_sample_logprob")
    torch.cuda.synchronize()
    assert torch.equal(act, act2) and torch.equal(sig, sig2) and torch.equal(logp, logp2)
    dist = torch.distributions.Normal(mu, actor.log_std.data.exp().expand_as(mu))
    assert torch.allclose(logp, dist.log_prob(act).sum(-1), rtol=1e-5, atol=1e-5)
    assert torch.allclose(sig, dist.stddev)


def test_encoder_parameter_write_reaches_the_fused_kernel():
    """The fused kernel reads the encoder through a packed table: a write to encoder.mean must invalidate it."""
    actor = build(48, (256, 256), 12, 5, "amp", seed=11)
    obs = torch.randn(64, 48, generator=torch.Generator().manual_seed(2)).cuda()
    with torch.inference_mode():
        mu0, _ = actor(obs)
        mu0 = mu0.clone()
    with torch.no_grad():
        actor.encoder.mean.add_(0.37)
    with torch.inference_mode():
        mu1, _ = actor(obs)
    mu_f, mu_l = both_paths(actor, obs)
    assert not torch.equal(mu0, mu1)
    assert torch.equal(mu1, mu_l) and torch.equal(mu_f, mu_l)


def test_unsupported_time_axis_falls_back_to_the_layered_kernels():
    actor = build(9, (40, 16), 3, 20, "amp", seed=1)         # T = 20 > 16
    assert not actor.engine.act_supported(50)
    obs = torch.randn(50, 9, generator=torch.Generator().manual_seed(3)).cuda()
    with torch.inference_mode():
        mu, _ = actor(obs)
    sd = {"actor." + k: v.detach().cpu() for k, v in actor.state_dict().items()}
    p = O.ActorParams.from_state_dict(sd, spike_ts=20, enc_eps=0.0, dec_tanh=False)
    mu_ref, _ = O.actor_forward(obs.cpu(), p)
    assert ((mu.cpu() - mu_ref).abs() <= REL_TOL * mu_ref.abs().max()).all()
