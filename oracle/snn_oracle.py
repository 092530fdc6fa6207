"""CPU oracle for the population-coded spiking actor-critic hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``fullysnnforleggedrobots_b200/`` may import
this file; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker / baseline.

This is a plain PyTorch-fp32 *restatement* (functional, no nn.Module, written from the
algorithm, not from the source text) of the reference's spiking actor:

    AMP/AMP_for_hardware-main/rsl_rl/rsl_rl/modules/actor_critic.py   (cited as AMP:<line>)

and its three sibling copies (RMA / CASSIE / HUMANOID; SURVEY.md §2.1 lists the deltas,
which are the ``enc_eps`` and ``dec_tanh`` switches below).

Parity pinning: the reference has NO tests / golden vectors for this path,
so the oracle is pinned against outputs of the reference module itself, imported from
/root/reference by ``tests/golden/make_golden.py`` (committed together with the vectors it
produced).  ``tests/test_oracle_golden.py`` checks every function here against them.

Tensor conventions (the layouts the CUDA path uses, not the reference's):
    obs            [B, O]            fp32
    enc spikes     [T, B, O*P]       {0,1}   (reference: [B, O*P, T], AMP:112)
    layer traces   v[l], s[l]        [T, B, N_l]
    mu             [B, A]
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch

# AMP:39-43
ENC_VTH = 0.999
VTH = 0.5
CDECAY = 0.5
VDECAY = 0.75
GRAD_WIN = 0.5


@dataclass
class ActorParams:
    """Flat view of a PopSpikeActor's parameters (state-dict names in comments)."""
    enc_mean: torch.Tensor        # actor.encoder.mean            [1, O, P]
    enc_std: torch.Tensor         # actor.encoder.std             [1, O, P]
    weights: List[torch.Tensor]   # actor.snn.hidden_layers.i.weight ..., actor.snn.out_pop_layer.weight  [N, K]
    biases: List[torch.Tensor]    # same, .bias                   [N]
    dec_w: torch.Tensor           # actor.decoder.decoder.weight  [A, 1, P]
    dec_b: torch.Tensor           # actor.decoder.decoder.bias    [A]
    log_std: torch.Tensor         # actor.log_std                 [A]
    spike_ts: int = 5
    enc_eps: float = 0.0          # HUMANOID: 1e-5 added to std^2 (HUM actor_critic.py:107)
    dec_tanh: bool = False        # CASSIE: tanh applied on decoder output (CAS actor_critic.py:147)

    @staticmethod
    def from_state_dict(sd: Dict[str, torch.Tensor], prefix: str = "actor.", **kw) -> "ActorParams":
        g = lambda k: sd[prefix + k].detach().to(torch.float32).cpu()
        ws, bs = [], []
        i = 0
        while f"{prefix}snn.hidden_layers.{i}.weight" in sd:
            ws.append(g(f"snn.hidden_layers.{i}.weight"))
            bs.append(g(f"snn.hidden_layers.{i}.bias"))
            i += 1
        ws.append(g("snn.out_pop_layer.weight"))
        bs.append(g("snn.out_pop_layer.bias"))
        return ActorParams(g("encoder.mean"), g("encoder.std"), ws, bs,
                           g("decoder.decoder.weight"), g("decoder.decoder.bias"), g("log_std"), **kw)


@dataclass
class Traces:
    pop_act: torch.Tensor = None                 # [B, K0] gaussian receptive-field activation
    enc_spikes: torch.Tensor = None              # [T, B, K0]
    volt: List[torch.Tensor] = field(default_factory=list)    # per layer [T, B, N]
    spikes: List[torch.Tensor] = field(default_factory=list)  # per layer [T, B, N]
    out_act: torch.Tensor = None                 # [B, A*P] spike rate of the output population
    pre_tanh: torch.Tensor = None                # [B, A] decoder output before the optional tanh


# --------------------------------------------------------------------------- forward

def encode(obs: torch.Tensor, p: ActorParams) -> Tuple[torch.Tensor, torch.Tensor]:
    """Gaussian population code + regular spike generator.  AMP:94-122.

    a = exp(-1/2 (x-mean)^2 / std^2)           AMP:105  (HUMANOID: / (std^2 + 1e-5))
    for t: u += a ; s_t = u > 0.999 ; u -= 0.999 s_t        AMP:116-119
    """
    B = obs.shape[0]
    x = obs.reshape(B, -1, 1)
    var = p.enc_std.pow(2)
    if p.enc_eps:
        var = var + p.enc_eps
    a = torch.exp(-0.5 * (x - p.enc_mean).pow(2) / var).reshape(B, -1)
    u = torch.zeros_like(a)
    out = []
    for _ in range(p.spike_ts):
        u = u + a
        s = (u > ENC_VTH).to(torch.float32)
        u = u - s * ENC_VTH
        out.append(s)
    return a, torch.stack(out, 0)


def lif_layers(enc_spikes: torch.Tensor, p: ActorParams, tr: Optional[Traces] = None) -> torch.Tensor:
    """LIF SpikeMLP, evaluated time-major exactly like the reference.  AMP:216-271.

    c <- 0.5 c + (W x + b)           AMP:228
    v <- 0.75 v (1 - s_prev) + c     AMP:230
    s <- v > 0.5                     AMP:231
    out_act = sum_t s_out / T        AMP:267,269
    """
    T, B = enc_spikes.shape[0], enc_spikes.shape[1]
    L = len(p.weights)
    dev = enc_spikes.device
    c = [torch.zeros(B, w.shape[0], device=dev) for w in p.weights]
    v = [torch.zeros(B, w.shape[0], device=dev) for w in p.weights]
    s = [torch.zeros(B, w.shape[0], device=dev) for w in p.weights]
    vt = [[] for _ in range(L)]
    st = [[] for _ in range(L)]
    acc = torch.zeros(B, p.weights[-1].shape[0], device=dev)
    for t in range(T):
        x = enc_spikes[t]
        for l in range(L):
            c[l] = c[l] * CDECAY + torch.addmm(p.biases[l], x, p.weights[l].t())
            v[l] = v[l] * VDECAY * (1.0 - s[l]) + c[l]
            s[l] = (v[l] > VTH).to(torch.float32)
            x = s[l]
            if tr is not None:
                vt[l].append(v[l])
                st[l].append(s[l])
        acc = acc + s[-1]
    if tr is not None:
        tr.volt = [torch.stack(q, 0) for q in vt]
        tr.spikes = [torch.stack(q, 0) for q in st]
    return acc / T


def decode(out_act: torch.Tensor, p: ActorParams, tr: Optional[Traces] = None) -> torch.Tensor:
    """Grouped conv1d(A, A, P, groups=A) == per-action dot over its population.  AMP:140-150."""
    A, _, P = p.dec_w.shape
    raw = (out_act.reshape(-1, A, P) * p.dec_w.reshape(1, A, P)).sum(-1) + p.dec_b
    if tr is not None:
        tr.pre_tanh = raw
    return torch.tanh(raw) if p.dec_tanh else raw


def actor_forward(obs: torch.Tensor, p: ActorParams, want_traces: bool = False):
    """PopSpikeActor.forward: (mu [B,A], std [A]).  AMP:300-320."""
    tr = Traces() if want_traces else None
    a, es = encode(obs, p)
    if tr is not None:
        tr.pop_act, tr.enc_spikes = a, es
    oa = lif_layers(es, p, tr)
    if tr is not None:
        tr.out_act = oa
    mu = decode(oa, p, tr)
    std = torch.exp(p.log_std)
    return (mu, std, tr) if want_traces else (mu, std)


# --------------------------------------------------------------------------- backward (BPTT)

def actor_backward(obs: torch.Tensor, p: ActorParams, tr: Traces, g_mu: torch.Tensor,
                   need_dobs: bool = False) -> Dict[str, torch.Tensor]:
    """Manual surrogate-gradient BPTT: what autograd does to the reference graph.

    PseudoSpikeRect.backward (AMP:161-167): ds/dv = 1[|v-0.5| < 0.5] = 1[0 < v < 1].
    PseudoEncoderSpikeRegular.backward (AMP:57-59): identity.
    Recurrences, t = T-1..0, l = L-1..0, zero boundary at t = T:
        g_s = g_up_t + g_v' * (-0.75 v_t)
        g_v = g_s * 1[0<v_t<1] + g_v' * 0.75 (1 - s_t)
        g_c = g_v + 0.5 g_c'
        dW += g_c^T x_t ; db += sum_b g_c ; g_down_t = g_c W
    """
    T = p.spike_ts
    L = len(p.weights)
    A, _, P = p.dec_w.shape
    B = obs.shape[0]
    G = g_mu
    if p.dec_tanh:
        G = G * (1.0 - torch.tanh(tr.pre_tanh).pow(2))
    grads: Dict[str, torch.Tensor] = {}
    oa = tr.out_act.reshape(B, A, P)
    grads["dec_w"] = (G.unsqueeze(-1) * oa).sum(0).reshape(A, 1, P)
    grads["dec_b"] = G.sum(0)
    # d out_act -> every step's output spike gets G*w/T
    g_up = [(G.unsqueeze(-1) * p.dec_w.reshape(1, A, P)).reshape(B, A * P) / T for _ in range(T)]
    dW = [torch.zeros_like(w) for w in p.weights]
    dB = [torch.zeros_like(b) for b in p.biases]
    for l in range(L - 1, -1, -1):
        v, s = tr.volt[l], tr.spikes[l]
        x_in = tr.enc_spikes if l == 0 else tr.spikes[l - 1]
        gv_n = torch.zeros_like(v[0])
        gc_n = torch.zeros_like(v[0])
        g_down = [None] * T
        for t in range(T - 1, -1, -1):
            g_s = g_up[t] + gv_n * (-VDECAY * v[t])
            win = ((v[t] - VTH).abs() < GRAD_WIN).to(torch.float32)
            g_v = g_s * win + gv_n * VDECAY * (1.0 - s[t])
            g_c = g_v + CDECAY * gc_n
            dW[l] += g_c.t() @ x_in[t]
            dB[l] += g_c.sum(0)
            g_down[t] = g_c @ p.weights[l]
            gv_n, gc_n = g_v, g_c
        g_up = g_down
    grads["weights"] = dW
    grads["biases"] = dB
    # Encoder (AMP:116-119 under autograd).  With u_t the potential after "+= a":
    #   s_t = STE(u_t) (identity grad) ; u_{t+1} = u_t - 0.999 s_t + a
    # so Gamma_t := dL/du_t = g_up_t + (1 - 0.999) Gamma_{t+1}, and dL/da = sum_t Gamma_t.
    gam = torch.zeros_like(g_up[0])
    g_a = torch.zeros_like(g_up[0])
    for t in range(T - 1, -1, -1):
        gam = g_up[t] + (1.0 - ENC_VTH) * gam
        g_a = g_a + gam
    O = obs.shape[1]
    x = obs.reshape(B, O, 1)
    var = p.enc_std.pow(2)
    if p.enc_eps:
        var = var + p.enc_eps
    d = x - p.enc_mean
    ga = (g_a * tr.pop_act).reshape(B, O, -1)
    grads["enc_mean"] = (ga * d / var).sum(0, keepdim=True)
    grads["enc_std"] = (ga * d.pow(2) / var.pow(2) * p.enc_std).sum(0, keepdim=True)
    if need_dobs:
        grads["obs"] = (ga * (-d) / var).sum(-1)
    return grads


# --------------------------------------------------------------------------- autograd twin
# Same forward, but differentiable with the reference's surrogate functions, so that the
# cpu_baseline "fwd+BPTT" leg exercises the same kind of autograd graph as the reference.

class _RectSpike(torch.autograd.Function):          # AMP:154-167
    @staticmethod
    def forward(ctx, v):
        ctx.save_for_backward(v)
        return (v > VTH).to(v.dtype)

    @staticmethod
    def backward(ctx, g):
        (v,) = ctx.saved_tensors
        return g * ((v - VTH).abs() < GRAD_WIN).to(g.dtype)


class _EncSpike(torch.autograd.Function):           # AMP:50-59
    @staticmethod
    def forward(ctx, u):
        return (u > ENC_VTH).to(u.dtype)

    @staticmethod
    def backward(ctx, g):
        return g.clone()


def actor_forward_autograd(obs: torch.Tensor, p: ActorParams):
    """Differentiable twin of actor_forward (time-major, one addmm per (layer, step))."""
    B = obs.shape[0]
    x = obs.reshape(B, -1, 1)
    var = p.enc_std.pow(2)
    if p.enc_eps:
        var = var + p.enc_eps
    a = torch.exp(-0.5 * (x - p.enc_mean).pow(2) / var).reshape(B, -1)
    u = torch.zeros_like(a)
    L = len(p.weights)
    dev = obs.device
    c = [torch.zeros(B, w.shape[0], device=dev) for w in p.weights]
    v = [torch.zeros(B, w.shape[0], device=dev) for w in p.weights]
    s = [torch.zeros(B, w.shape[0], device=dev) for w in p.weights]
    acc = torch.zeros(B, p.weights[-1].shape[0], device=dev)
    for _ in range(p.spike_ts):
        u = u + a
        es = _EncSpike.apply(u)
        u = u - es * ENC_VTH
        h = es
        for l in range(L):
            c[l] = c[l] * CDECAY + torch.addmm(p.biases[l], h, p.weights[l].t())
            v[l] = v[l] * VDECAY * (1.0 - s[l]) + c[l]
            s[l] = _RectSpike.apply(v[l])
            h = s[l]
        acc = acc + s[-1]
    mu = decode(acc / p.spike_ts, p)
    return mu, torch.exp(p.log_std)


# --------------------------------------------------------------------------- PPO pieces

LOG_SQRT_2PI = 0.5 * math.log(2.0 * math.pi)


def normal_log_prob(actions, mu, std):
    """torch.distributions.Normal.log_prob(...).sum(-1)   (AMP:420-421)."""
    return (-((actions - mu) ** 2) / (2.0 * std * std) - torch.log(std) - LOG_SQRT_2PI).sum(-1)


def normal_entropy(mu, std):
    """Normal.entropy().sum(-1)  (AMP:406-408)."""
    return (0.5 + LOG_SQRT_2PI + torch.log(std)).expand_as(mu).sum(-1)


def critic_forward(x: torch.Tensor, ws: List[torch.Tensor], bs: List[torch.Tensor]) -> torch.Tensor:
    """Dense ELU MLP critic, Linear->ELU->...->Linear(.,1).  AMP:359-368, 428-430."""
    h = x
    for i, (w, b) in enumerate(zip(ws, bs)):
        h = torch.addmm(b, h, w.t())
        if i + 1 < len(ws):
            h = torch.nn.functional.elu(h)
    return h


def gae_returns(rewards, dones, values, last_values, gamma, lam):
    """RolloutStorage.compute_returns.  AMP storage/rollout_storage.py:124-138.

    rewards/values [S, N, 1], dones [S, N, 1] (uint8/bool), last_values [N, 1].
    Returns (returns, normalised advantages).
    """
    S = rewards.shape[0]
    returns = torch.zeros_like(rewards)
    adv = torch.zeros_like(last_values)
    for t in range(S - 1, -1, -1):
        nv = last_values if t == S - 1 else values[t + 1]
        nt = 1.0 - dones[t].to(torch.float32)
        delta = rewards[t] + nt * gamma * nv - values[t]
        adv = delta + nt * gamma * lam * adv
        returns[t] = adv + values[t]
    a = returns - values
    a = (a - a.mean()) / (a.std() + 1e-8)
    return returns, a


def ppo_losses(logp, old_logp, adv, value, target_values, returns, entropy,
               clip=0.2, value_coef=1.0, ent_coef=0.0, clipped_value=True):
    """Clipped surrogate + clipped value loss + entropy bonus.  AMP algorithms/ppo.py:153-171."""
    ratio = torch.exp(logp - old_logp.squeeze(-1))
    a = adv.squeeze(-1)
    sur = torch.max(-a * ratio, -a * torch.clamp(ratio, 1.0 - clip, 1.0 + clip)).mean()
    if clipped_value:
        vc = target_values + (value - target_values).clamp(-clip, clip)
        vl = torch.max((value - returns).pow(2), (vc - returns).pow(2)).mean()
    else:
        vl = (returns - value).pow(2).mean()
    loss = sur + value_coef * vl - ent_coef * entropy.mean()
    return loss, sur, vl


def kl_gaussian(mu, sigma, old_mu, old_sigma):
    """Adaptive-LR KL estimate.  AMP algorithms/ppo.py:141-143."""
    return torch.sum(torch.log(sigma / old_sigma + 1.e-5)
                     + (old_sigma.pow(2) + (old_mu - mu).pow(2)) / (2.0 * sigma.pow(2)) - 0.5, dim=-1)
