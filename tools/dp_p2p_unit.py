"""Kernel-level check of snn_clip_adam_kl_p2p on 2..8 ranks (torchrun): random gradient buckets per rank, the summed
bucket and the updated parameters against NCCL all_reduce + snn_clip_adam_kl on copies of the same data, several steps
(both parities of the staging buffer, the adaptive-KL rule on the summed tail)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

from fullysnnforleggedrobots_b200 import _lib
from fullysnnforleggedrobots_b200.dist import P2PGradientSum

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
dist.init_process_group("nccl")
L = _lib.lib()
p = lambda t: C.c_void_p(t.data_ptr())
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
worst = 0.0
for n in (303_117, 1000, 5_000_001):
    ntot = n + 1
    p2p = P2PGradientSum(ntot)
    g0 = torch.Generator(device="cuda").manual_seed(7)
    prm = torch.randn(n, device="cuda", generator=g0)
    state = {}
    for mode in ("p2p", "nccl"):
        state[mode] = dict(prm=prm.clone(), m=torch.zeros(n, device="cuda"), v=torch.zeros(n, device="cuda"),
                           lr=torch.full((1,), 1e-3, device="cuda"), step=torch.zeros(1, dtype=torch.int32, device="cuda"),
                           scratch=torch.empty(4096, dtype=torch.uint8, device="cuda"))
    for it in range(4):
        gr = torch.Generator(device="cuda").manual_seed(1000 * it + rank)
        g = torch.randn(ntot, device="cuda", generator=gr) * 0.01
        g[n] = 0.004 * (1 + it)                      # KL statistic of this rank: the mean crosses the rule's thresholds
        ga, gb = g.clone(), g.clone()
        s = state["p2p"]
        _lib.check(L.snn_clip_adam_kl_p2p(p(s["prm"]), p(ga), p(s["m"]), p(s["v"]), n, ntot, p(s["lr"]), p(s["step"]), 0.9, 0.999,
                                          1e-8, 0.0, 1.0, 1.0 / world, n, 1.0 / world, 0.01, None, p(s["scratch"]), p2p.stage,
                                          p2p.flags, p2p.rank, p2p.world_size, p2p.stride, p(p2p.epoch), st()), "p2p")
        s = state["nccl"]
        dist.all_reduce(gb)
        _lib.check(L.snn_clip_adam_kl(p(s["prm"]), p(gb), p(s["m"]), p(s["v"]), n, p(s["lr"]), p(s["step"]), 0.9, 0.999, 1e-8, 0.0,
                                      1.0, 1.0 / world, p(gb[n:]), 1.0 / world, 0.01, None, p(s["scratch"]), st()), "nccl+adam")
        torch.cuda.synchronize()
        dg = float((ga - gb).abs().max() / gb.abs().max())
        dp = float((state["p2p"]["prm"] - state["nccl"]["prm"]).abs().max())
        lr_eq = bool(torch.equal(state["p2p"]["lr"], state["nccl"]["lr"]))
        allg = [torch.zeros_like(ga[:64]) for _ in range(world)]
        dist.all_gather(allg, ga[:64].contiguous())
        same = all(torch.equal(a, allg[0]) for a in allg)
        worst = max(worst, dg)
        if rank == 0:
            print(f"n={n} step {it}: summed bucket p2p vs nccl max rel diff {dg:.2e}, params max |diff| {dp:.2e}, lr equal {lr_eq} "
                  f"({float(state['p2p']['lr']):.2e}), ranks identical {same}", flush=True)
    p2p.close()
if rank == 0:
    print("OK" if worst < 1e-6 else "MISMATCH", flush=True)
dist.barrier()
torch.cuda.synchronize()
os._exit(0)
