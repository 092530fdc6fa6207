"""Inference forward of the spiking actor: the one-launch fused kernel (snn_fwd_act) against the layered path
(encode -> one scan-GEMM per layer -> decode), CUDA events over REP back-to-back calls, C1 dims."""
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from fullysnnforleggedrobots_b200 import PopSpikeActor

REP = 200
sizes = [int(a) for a in sys.argv[1:]] or [16, 256, 1024, 4096, 16384, 65536]
torch.manual_seed(0)
actor = PopSpikeActor(48, 12, 10, 10, [256, 256], (-5, 5), math.sqrt(0.15), spike_ts=5).cuda()
layers = actor.snn.linear_layers()
params = (actor.encoder.mean.data, actor.encoder.std.data, [l.weight.data for l in layers], [l.bias.data for l in layers],
          actor.decoder.decoder.weight.data, actor.decoder.decoder.bias.data)
eng = actor.engine
flops_per_sample = 2 * 5 * (480 * 256 + 256 * 256 + 256 * 120)
for B in sizes:
    obs = torch.randn(B, 48, device="cuda")
    mu = torch.empty(B, 12, device="cuda")
    row = {"B": B}
    for name, fused in (("fused", True), ("layered", False)):
        eng.fused_infer = fused

        def call():
            if fused:       # the one-launch kernel whatever the batch size (forward() would pick by itself)
                eng.act(obs, *params, None, None, mu, None, None, None)
            else:
                eng.forward(obs, *params, train=False, out_mu=mu)
        for _ in range(5):
            call()
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(REP):
                call()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / REP)
        row[name + "_us"] = round(best * 1e3, 2)
        row[name + "_Msteps_s"] = round(B / best / 1e3, 2)
        row[name + "_algo_tflops"] = round(B * flops_per_sample / (best * 1e-3) / 1e12, 1)
    print(json.dumps(row))
