"""Print a full parity table (GPU kernels vs CPU oracle) for the golden cases.  Debug aid."""
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)

import torch

from cases import CASES, make_gmu, make_obs, make_state_dict
from fullysnnforleggedrobots_b200 import PopSpikeActor
from helpers import mismatch_rate, oracle_params, rel_err
from oracle import snn_oracle as O


def main():
    names = sys.argv[1:]
    for c in CASES:
        if names and c.name not in names:
            continue
        print(f"=== {c.name}: O={c.obs_dim} H={c.hidden} A={c.act_dim} B={c.batch} T={c.spike_ts}", flush=True)
        actor = PopSpikeActor(c.obs_dim, c.act_dim, c.pop, c.pop, list(c.hidden), (-5, 5), math.sqrt(0.15),
                              spike_ts=c.spike_ts, enc_eps=c.enc_eps, dec_tanh=c.dec_tanh)
        actor.load_state_dict({k[len("actor."):]: v for k, v in make_state_dict(c).items()})
        actor = actor.cuda()
        p = oracle_params(c)
        obs_cpu = make_obs(c)
        G = make_gmu(c)
        mu_ref, _, tr_ref = O.actor_forward(obs_cpu, p, want_traces=True)
        gr = O.actor_backward(obs_cpu, p, tr_ref, G, need_dobs=True)
        layers = actor.snn.linear_layers()
        try:
            mu, trace = actor.engine.forward(obs_cpu.cuda(), actor.encoder.mean.data, actor.encoder.std.data,
                                             [l.weight.data for l in layers], [l.bias.data for l in layers],
                                             actor.decoder.decoder.weight.data, actor.decoder.decoder.bias.data,
                                             train=True)
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            print("  forward FAILED:", e)
            return 1
        tr = actor.engine.unpack_trace(trace, c.batch)
        print(f"  enc spikes mismatch {mismatch_rate(tr['enc_spikes'], tr_ref.enc_spikes):.2e}"
              f" (rate {tr_ref.enc_spikes.mean():.3f})")
        for l in range(len(p.weights)):
            same = tr["spikes"][l] == tr_ref.spikes[l]
            dv = ((tr["volt"][l] - tr_ref.volt[l]).abs() * same).max().item()
            dv0 = (tr["volt"][l][0] - tr_ref.volt[l][0]).abs().max().item()
            print(f"  layer {l}: spike mismatch {mismatch_rate(tr['spikes'][l], tr_ref.spikes[l]):.2e} "
                  f"(rate {tr_ref.spikes[l].mean():.3f})  max|dv| on matching {dv:.2e}  t0 {dv0:.2e} "
                  f"(|v|max {tr_ref.volt[l].abs().max():.2f})")
        print(f"  mu rel err {rel_err(mu.cpu(), mu_ref):.2e}")
        try:
            obs = obs_cpu.cuda().requires_grad_(True)
            mu2, _ = actor(obs)
            (mu2 * G.cuda()).sum().backward()
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            print("  backward FAILED:", e)
            return 1
        for l in range(len(layers) - 1, -1, -1):
            print(f"  dW{l} {rel_err(layers[l].weight.grad.cpu(), gr['weights'][l]):.2e}   "
                  f"db{l} {rel_err(layers[l].bias.grad.cpu(), gr['biases'][l]):.2e}")
        print(f"  dec_w {rel_err(actor.decoder.decoder.weight.grad.cpu(), gr['dec_w']):.2e}  "
              f"dec_b {rel_err(actor.decoder.decoder.bias.grad.cpu(), gr['dec_b']):.2e}  "
              f"enc_mean {rel_err(actor.encoder.mean.grad.cpu(), gr['enc_mean']):.2e}  "
              f"enc_std {rel_err(actor.encoder.std.grad.cpu(), gr['enc_std']):.2e}  "
              f"d_obs {rel_err(obs.grad.cpu(), gr['obs']):.2e}", flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
