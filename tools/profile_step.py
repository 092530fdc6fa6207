"""One training forward + BPTT backward of the C1-dims actor at a PPO minibatch (for ncu captures)."""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from fullysnnforleggedrobots_b200 import PopSpikeActor

B = int(sys.argv[1]) if len(sys.argv) > 1 else 24576
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
torch.manual_seed(0)
actor = PopSpikeActor(48, 12, 10, 10, [256, 256], (-5, 5), math.sqrt(0.15), spike_ts=5).cuda()
obs = torch.randn(B, 48, device="cuda")
G = torch.randn(B, 12, device="cuda")
for _ in range(reps):
    for p in actor.parameters():
        p.grad = None
    mu, std = actor(obs)
    (mu * G).sum().backward()
torch.cuda.synchronize()
print("ok", float(mu.abs().mean()))
