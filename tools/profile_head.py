"""Decoder + PPO actor head + output-layer scan (snn_bwd_ppo_top = actor_head_kernel + top_scan_rows_kernel) at a PPO minibatch:
CUDA-event time per launch, or one launch for an ncu capture (argv[2] = 1)."""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from fullysnnforleggedrobots_b200 import PopSpikeActor

B = int(sys.argv[1]) if len(sys.argv) > 1 else 24576
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
torch.manual_seed(0)
A = 12
actor = PopSpikeActor(48, A, 10, 10, [256, 256], (-5, 5), math.sqrt(0.15), spike_ts=5).cuda()
eng = actor.engine
layers = actor.snn.linear_layers()
params = (actor.encoder.mean.data, actor.encoder.std.data, [l.weight.data for l in layers], [l.bias.data for l in layers],
          actor.decoder.decoder.weight.data, actor.decoder.decoder.bias.data)
obs = torch.randn(B, 48, device="cuda")
mu, trace = eng.forward(obs, *params, train=True)
log_std = torch.zeros(A, device="cuda")
old_mu, old_sigma = mu + 0.1 * torch.randn_like(mu), torch.ones(B, A, device="cuda")
actions = old_mu + torch.randn_like(mu)
old_logp = torch.distributions.Normal(old_mu, old_sigma).log_prob(actions).sum(-1)
adv = torch.randn(B, device="cuda")
part = torch.empty(eng.ppo_head_part_bytes(B), dtype=torch.uint8, device="cuda")
ws = torch.empty(eng.workspace_bytes(B, True), dtype=torch.uint8, device="cuda")
mu_o, g_mu_o = torch.empty_like(mu), torch.empty_like(mu)


def run():
    eng.backward_ppo_top(obs, trace, params, log_std, actions, old_logp, adv, old_mu, old_sigma, 0.2, part, ws, mu_o, g_mu_o)


run()
torch.cuda.synchronize()
if reps > 1:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    print(f"snn_bwd_ppo_top B={B}: {e0.elapsed_time(e1) / reps * 1e3:.1f} us per launch")
