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
di = buf.cpu().view(16, 148, 4, 4)      # int64: the %globaltimer stamps need all 64 bits
d = di.double()
# wait sites of nm_gemm.cuh (fwd / dx) — the dW kernel reuses the same slots with its own meaning (dw_gemm.cuh)
names = {0: ("producer", ["w_empty", "g_empty(dx)", "-"]), 1: ("mma", ["acc_empty", "w_full", "s_full"]),
         2: ("epilogue", ["acc_full", "-", "-"]), 3: ("expander", ["bits_full", "slot_free", "-"])}
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
    if slot < 4 or slot >= 8:
        # nm_gemm kernels: %globaltimer stamps of every CTA (epilogue warp 4): first / last instruction
        st, en = di[slot][:, 2, 1], di[slot][:, 2, 2]
        live = st > 0
        if live.any():
            t0 = st[live].min()
            st, en = (st[live] - t0).double(), (en[live] - t0).double()
            t0 = 0.0
            print(f"   wall clock (us): CTA starts 0.0 .. {(st.max() - t0).item() / 1e3:.1f}, ends {(en.min() - t0).item() / 1e3:.1f} .. "
                  f"{(en.max() - t0).item() / 1e3:.1f}; mean CTA life {(en - st).mean().item() / 1e3:.1f}; "
                  f"cycles/ns of the epilogue warp {(x[:, 2, 3][live] / (en - st)).mean().item():.3f}")

# bubbles between dependent kernels: last CTA end of one kernel -> first CTA start of the next (absolute %globaltimer)
rows = []
for slot in range(16):
    if slot >= 4 and slot < 8:
        continue
    st, en = di[slot][:, 2, 1], di[slot][:, 2, 2]
    live = st > 0
    if live.any():
        rows.append((int(st[live].min()), int(st[live].max()), int(en[live].min()), int(en[live].max()), slot))
rows.sort()
if rows:
    t0 = rows[0][0]
    print("kernel chain (us since the first CTA of the first nm_gemm kernel): first start, last start, first end, last end")
    prev_end = None
    for a, b, c, e, slot in rows:
        kind = "fwd" if slot < 4 else "dx"
        gap = "" if prev_end is None else f"   bubble since the previous kernel's last CTA: {(a - prev_end) / 1e3:.1f} us"
        print(f"  {kind} layer {slot % 4 if slot < 8 else slot % 8}: {(a - t0) / 1e3:8.1f} {(b - t0) / 1e3:8.1f} {(c - t0) / 1e3:8.1f} {(e - t0) / 1e3:8.1f}{gap}")
        prev_end = e
