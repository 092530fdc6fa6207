"""The dense ELU critic (48 -> 256 -> 256 -> 1) forward + backward on the tensor-core MLP kernels at a PPO minibatch:
CUDA-event time per forward / backward, or a single pass for an ncu capture (argv[2] = 1)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from fullysnnforleggedrobots_b200.engine import CriticEngine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 24576
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
torch.manual_seed(0)
seq = torch.nn.Sequential(torch.nn.Linear(48, 256), torch.nn.ELU(), torch.nn.Linear(256, 256), torch.nn.ELU(),
                          torch.nn.Linear(256, 1)).cuda()
ce = CriticEngine.from_sequential(seq)
ce.repack()
x = torch.randn(B, 48, device="cuda")
gv = torch.randn(B, device="cuda")
for p in seq.parameters():
    p.grad = torch.zeros_like(p)
grads = ([l.weight.grad for l in ce.linears], [l.bias.grad for l in ce.linears])


def fwd():
    return ce.forward(x)


def bwd(tr):
    ce.backward(gv, tr, B, out=grads)


v, tr = fwd()
bwd(tr)
torch.cuda.synchronize()
if reps > 1:
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    for _ in range(reps):
        v, tr = fwd()
    e[1].record()
    for _ in range(reps):
        bwd(tr)
    e[2].record()
    torch.cuda.synchronize()
    print(f"critic B={B}: forward {e[0].elapsed_time(e[1]) / reps * 1e3:.1f} us, backward {e[1].elapsed_time(e[2]) / reps * 1e3:.1f} us")
