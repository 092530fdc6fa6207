"""Inside the fused rollout kernel (act_fused.cuh): where do the warp roles wait, and when does each layer finish?
-DSNN_ABLATE tools build only.  usage: python tools/act_cycles.py [B]"""
import ctypes as C
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

os.environ["SNN_LIB"] = "ablate"
from fullysnnforleggedrobots_b200 import _lib as _l
_l.build(ablate=True)
from fullysnnforleggedrobots_b200 import PopSpikeActor, _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
torch.manual_seed(0)
actor = PopSpikeActor(48, 12, 10, 10, [256, 256], (-5, 5), math.sqrt(0.15), spike_ts=5).cuda()
obs = torch.randn(B, 48, device="cuda")
L = _lib.lib()
buf = torch.zeros(148 * 32, dtype=torch.int64, device="cuda")
with torch.inference_mode():
    for _ in range(3):
        actor(obs)
    torch.cuda.synchronize()
    L.snn_debug_set_cycle_buffer(C.c_void_p(buf.data_ptr()))
    actor(obs)
    torch.cuda.synchronize()
    L.snn_debug_set_cycle_buffer(None)
d = buf.cpu().view(148, 32).double()
live = d[:, 3] > 0
d = d[live]
print(f"B={B}: {int(live.sum())} CTAs; kilo-cycles, mean over CTAs (max)")
roles = [("producer", ["w_empty"]), ("mma", ["layer_bar", "w_full", "e_full"]), ("chunks", ["e_empty", "prev_tile"]),
         ("epilogue0", ["acc_full"]), ("epilogue1", ["acc_full"])]
for r, (nm, ws) in enumerate(roles):
    x = d[:, 4 * r: 4 * r + 4]
    parts = ", ".join(f"wait {w} {x[:, i].mean() / 1e3:6.1f}" for i, w in enumerate(ws))
    print(f"  {nm:10s} total {x[:, 3].mean() / 1e3:6.1f} ({x[:, 3].max() / 1e3:6.1f}) | {parts}")
st = d[:, 20:32]
names = [f"layer {l} accumulators complete" for l in range(4)] + [f"layer {l} MMAs begin" for l in range(5)] + ["encode step begins", "encode step done", "-"]
print("fused path:", actor.engine.act_supported(B), "launches", L.snn_launch_count())
for k in range(12):
    if st[:, k].max() > 0:
        print(f"  stamp {names[k]:26s} {st[:, k].mean() / 1e3:6.1f} ({st[:, k].max() / 1e3:6.1f})")
