"""Per-kernel, per-layer device times of one training forward + backward (CUDA events inside the library)."""
import ctypes as C
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

# ablation switches / cycle counters live in the -DSNN_ABLATE tools build only (never in the release library)
if os.environ.get("SNN_DBG", "0") != "0":
    os.environ["SNN_LIB"] = "ablate"
    from fullysnnforleggedrobots_b200 import _lib as _l
    _l.build(ablate=True)
from fullysnnforleggedrobots_b200 import PopSpikeActor, _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 24576
dims = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [48, 256, 256, 12]
T = int(sys.argv[3]) if len(sys.argv) > 3 else 5
PL = int(sys.argv[4]) if len(sys.argv) > 4 else 3
reps = 5
torch.manual_seed(0)
actor = PopSpikeActor(dims[0], dims[-1], 10, 10, dims[1:-1], (-5, 5), math.sqrt(0.15), spike_ts=T, fwd_planes=PL).cuda()
obs = torch.randn(B, dims[0], device="cuda")
G = torch.randn(B, dims[-1], device="cuda")
L = _lib.lib()


def step():
    for p in actor.parameters():
        p.grad = None
    mu, _ = actor(obs)
    (mu * G).sum().backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
L.snn_profile_enable(1)
for _ in range(reps):
    step()
ms, fl, ln = (C.c_double * 32)(), (C.c_double * 32)(), (C.c_int64 * 32)()
L.snn_profile_read(ms, fl, ln)
L.snn_profile_enable(0)
names = ["fwd", "dx", "dw", "other"]
print(f"B={B} dims={dims} T={T} fwd_planes={PL} SNN_DBG={os.environ.get('SNN_DBG', '0')}")
tot = 0.0
for l in range(8):
    for k in range(4):
        i = k + 4 * l
        if ln[i]:
            us = ms[i] / ln[i] * 1e3
            tot += ms[i] / reps * 1e3
            tf = fl[i] / (ms[i] * 1e-3) / 1e12 if ms[i] > 0 and fl[i] > 0 else 0
            print(f"  layer {l} {names[k]:5s}: {us:8.1f} us/launch x{ln[i] // reps}  algo {tf:7.1f} TFLOP/s")
print(f"  total per fwd+bwd: {tot:.1f} us")
