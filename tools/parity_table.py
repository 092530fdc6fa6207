"""Parity table: CUDA path (through the C ABI) vs the CPU oracle, on the golden cases and at rollout / minibatch size.

Columns: spike-train mismatch rate (encoder and worst LIF layer; north-star bar 1e-3), max |dv| on matching spikes,
mu and worst-gradient error relative to the tensor's max-abs (bar 1e-4).   python tools/parity_table.py [out.md]
"""
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)

import torch

from cases import CASES, Case, make_gmu, make_obs, make_state_dict
from fullysnnforleggedrobots_b200 import PopSpikeActor
from helpers import mismatch_rate, oracle_params, rel_err
from oracle import snn_oracle as O

BIG = [Case("c1_rollout_4096", "amp", 48, (256, 256), 12, 4096, seed=31, full=False),
       Case("c1_minibatch_24576", "amp", 48, (256, 256), 12, 24576, seed=32, full=False, trained=True),
       Case("humanoid_3072_shard", "humanoid", 38, (256, 256, 256), 10, 3072, seed=33, full=False),
       Case("cassie_latent_T8_2048", "cassie", 187, (512, 256, 128), 12, 2048, spike_ts=8, seed=34, full=False)]


def run(c):
    actor = PopSpikeActor(c.obs_dim, c.act_dim, c.pop, c.pop, list(c.hidden), (-5, 5), math.sqrt(0.15),
                          spike_ts=c.spike_ts, enc_eps=c.enc_eps, dec_tanh=c.dec_tanh)
    actor.load_state_dict({k[len("actor."):]: v for k, v in make_state_dict(c).items()})
    actor = actor.cuda()
    p = oracle_params(c)
    obs_cpu, G = make_obs(c), make_gmu(c)
    t0 = time.time()
    mu_ref, _, tr_ref = O.actor_forward(obs_cpu, p, want_traces=True)
    gr = O.actor_backward(obs_cpu, p, tr_ref, G, need_dobs=True)
    t_cpu = time.time() - t0
    layers = actor.snn.linear_layers()
    mu, trace = actor.engine.forward(obs_cpu.cuda(), actor.encoder.mean.data, actor.encoder.std.data,
                                     [l.weight.data for l in layers], [l.bias.data for l in layers],
                                     actor.decoder.decoder.weight.data, actor.decoder.decoder.bias.data, train=True)
    tr = actor.engine.unpack_trace(trace, c.batch)
    enc_mm = mismatch_rate(tr["enc_spikes"], tr_ref.enc_spikes)
    B = c.batch
    # rows (samples) on which every spike and every surrogate window 1[0 < v < 1] agrees with the oracle: a tie at the
    # spike threshold or at a window edge (|v - edge| below the fp32 summation-order noise) legitimately changes that
    # row's action / gradient; all other rows must agree to 1e-4
    clean = torch.ones(B, dtype=torch.bool)

    def rows_of(x):          # [T, B, N] or [B, N, T] style traces -> bool [B]: any mismatch in the row
        x = x if x.shape[1] == B else x.transpose(0, 1)
        return x.transpose(0, 1).reshape(B, -1).any(dim=1)

    clean &= ~rows_of(tr["enc_spikes"] != tr_ref.enc_spikes)
    lay_mm, win_mm, dvs = 0.0, 0.0, []
    for l in range(len(p.weights)):
        ds = tr["spikes"][l] != tr_ref.spikes[l]
        wg = (tr["volt"][l] > 0) & (tr["volt"][l] < 1)
        wr = (tr_ref.volt[l] > 0) & (tr_ref.volt[l] < 1)
        lay_mm = max(lay_mm, float(ds.double().mean()))
        win_mm = max(win_mm, float((wg != wr).double().mean()))
        clean &= ~rows_of(ds) & ~rows_of(wg != wr)
        dvs.append((tr["volt"][l] - tr_ref.volt[l]).abs())
    dv = max(float(x[:, clean].max()) if clean.any() else float("nan") for x in dvs)
    obs = obs_cpu.cuda().requires_grad_(True)
    mu2, _ = actor(obs)
    (mu2 * G.cuda()).sum().backward()
    torch.cuda.synchronize()
    errs = {"dW": max(rel_err(layers[l].weight.grad.cpu(), gr["weights"][l]) for l in range(len(layers))),
            "db": max(rel_err(layers[l].bias.grad.cpu(), gr["biases"][l]) for l in range(len(layers))),
            "dec": max(rel_err(actor.decoder.decoder.weight.grad.cpu(), gr["dec_w"]),
                       rel_err(actor.decoder.decoder.bias.grad.cpu(), gr["dec_b"])),
            "enc": max(rel_err(actor.encoder.mean.grad.cpu(), gr["enc_mean"]),
                       rel_err(actor.encoder.std.grad.cpu(), gr["enc_std"]))}
    worst = max(errs, key=errs.get)
    scale_mu, scale_do = float(mu_ref.abs().max()), float(gr["obs"].abs().max())
    e_mu = float((mu.cpu() - mu_ref)[clean].abs().max() / scale_mu) if clean.any() else float("nan")
    e_do = float((obs.grad.cpu() - gr["obs"])[clean].abs().max() / scale_do) if clean.any() else float("nan")
    return (f"| {c.name} | {c.variant} | {c.obs_dim}→{'-'.join(map(str, c.hidden))}→{c.act_dim} | {c.batch} | {c.spike_ts} | "
            f"{enc_mm:.1e} | {lay_mm:.1e} | {win_mm:.1e} | {100.0 * float((~clean).double().mean()):.3f} % | {dv:.1e} | "
            f"{e_mu:.1e} | {e_do:.1e} | {errs[worst]:.1e} ({worst}) | {t_cpu:.1f} |")


def self_noise(c):
    """The oracle against ITSELF with the hidden units of every hidden layer permuted (rows of W_l, bias_l and the
    matching columns of W_{l+1}): the same network, another fp32 summation order in every layer above the first.
    Shows how many threshold ties two correct fp32 implementations produce at this batch size."""
    import dataclasses
    p = oracle_params(c)
    obs = make_obs(c)
    mu_a, _, tr_a = O.actor_forward(obs, p, want_traces=True)
    g = torch.Generator().manual_seed(123)
    ws, bs = [w.clone() for w in p.weights], [b.clone() for b in p.biases]
    for l in range(len(ws) - 1):
        perm = torch.randperm(ws[l].shape[0], generator=g)
        ws[l], bs[l] = ws[l][perm].contiguous(), bs[l][perm].contiguous()
        ws[l + 1] = ws[l + 1][:, perm].contiguous()
    mu_b, _, tr_b = O.actor_forward(obs, dataclasses.replace(p, weights=ws, biases=bs), want_traces=True)
    top = len(ws) - 1
    ds = tr_a.spikes[top] != tr_b.spikes[top]
    rows = ds.transpose(0, 1).reshape(c.batch, -1).any(dim=1)
    bad_mu = ((mu_a - mu_b).abs().max(dim=1).values > 1e-4 * float(mu_a.abs().max()))
    return (f"oracle vs oracle with permuted hidden units, {c.name}: output-layer spike mismatch "
            f"{float(ds.double().mean()):.1e}, rows with a flipped output spike {100.0 * float(rows.double().mean()):.3f} %, "
            f"rows whose action differs by more than 1e-4: {100.0 * float(bad_mu.double().mean()):.3f} %")


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else None
    lines = ["# Parity of the CUDA path against the CPU oracle (fp32-parity mode; oracle pinned to the reference fixtures)", "",
             "Bars (BASELINE.json north_star): spike mismatch <= 1e-3, actions / gradients within 1e-4 relative "
             "(relative to the tensor's max-abs).  `tools/parity_table.py` on one B200.", "",
             "Rows with a tie = samples where a membrane value sits within the fp32 summation-order noise of the spike "
             "threshold (0.5) or of a surrogate-window edge (0 or 1), so that the two implementations legitimately decide "
             "differently; the action / d obs columns are over all OTHER rows, the parameter gradients (batch sums) over all rows.", "",
             "| case | fork | dims | B | T | encoder spike mismatch | worst LIF-layer spike mismatch | worst surrogate-window "
             "mismatch | rows with a tie | max abs dv (tie-free rows) | mu rel err (tie-free rows) | d obs rel err (tie-free rows) | "
             "worst parameter-gradient rel err | oracle s |", "|---|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|"]
    for c in list(CASES) + BIG:
        lines.append(run(c))
        print(lines[-1], flush=True)
    lines += ["", "Reference self-noise (CPU only, no kernel involved):", ""]
    for c in BIG[:2]:
        lines.append("* " + self_noise(c))
        print(lines[-1], flush=True)
    text = "\n".join(lines) + "\n"
    if out:
        open(out, "w").write(text)


if __name__ == "__main__":
    main()
