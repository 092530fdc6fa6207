"""Where do the warp roles of the scan-GEMM kernels wait?  Cycle counters of every mbarrier wait site."""
import ctypes as C
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

# ablation switches / cycle counters live in the -DSNN_ABLATE tools build only (never in the release library)
if True:
    os.environ["SNN_LIB"] = "ablate"
    from fullysnnforleggedrobots_b200 import _lib as _l
    _l.build(ablate=True)
from fullysnnforleggedrobots_b200 import PopSpikeActor, _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 24576
torch.manual_seed(0)
actor = PopSpikeActor(48, 12, 10, 10, [256, 256], (-5, 5), math.sqrt(0.15), spike_ts=5).cuda()
obs = torch.randn(B, 48, device="cuda")
G = torch.randn(B, 12, device="cuda")
L = _lib.lib()
buf = torch.zeros(16 * 148 * 16, dtype=torch.int64, device="cuda")


def step():
    for p in actor.parameters():
        p.grad = None
    mu, _ = actor(obs)
    (mu * G).sum().backward()


for _ in range(2):
    step()
torch.cuda.synchronize()
L.snn_debug_set_cycle_buffer(C.c_void_p(buf.data_ptr()))
buf.zero_()
step()
torch.cuda.synchronize()
L.snn_debug_set_cycle_buffer(None)
d = buf.cpu().view(16, 148, 4, 4).double()
# wait sites of nm_gemm.cuh (fwd / dx) — the dW kernel reuses the same slots with its own meaning (dw_gemm.cuh)
names = {0: ("producer", ["w_empty", "g_empty(dx)", "-"]), 1: ("mma", ["acc_empty", "w_full", "s_full"]),
         2: ("epilogue", ["acc_full", "-", "-"]), 3: ("expander", ["bits_full", "staging_free", "-"])}
print(f"SNN_DBG={os.environ.get('SNN_DBG', '0')}  (mean over CTAs, kilo-cycles)")
for slot in range(16):
    x = d[slot]
    if x.sum() == 0:
        continue
    kind = "fwd" if slot < 4 else ("dw" if slot < 8 else "dx")
    print(f" {kind} layer {slot % 4 if slot < 8 else slot % 8}:")
    for r in range(4):
        tot = x[:, r, 3].mean().item() / 1e3
        if tot == 0:
            continue
        nm, ws = names[r]
        parts = ", ".join(f"wait {w} {x[:, r, i].mean().item() / 1e3:7.1f}" for i, w in enumerate(ws) if w != "-")
        print(f"   {nm:9s} total {tot:7.1f} | {parts}")
