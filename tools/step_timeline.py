"""Kernel timeline of the PPO update (bench.py's workload) from CUPTI via torch.profiler: per-kernel device time
inside the replayed CUDA graphs, stream occupancy, and the gaps between kernels on the actor's stream.

    python tools/step_timeline.py [eager|graph] [out.md]
"""
import os
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import ProfilerActivity, profile

from fullysnnforleggedrobots_b200 import PPO, ActorCritic

mode = sys.argv[1] if len(sys.argv) > 1 else "graph"
out = sys.argv[2] if len(sys.argv) > 2 else None
N, S, O, A = 4096, 24, 48, 12
dev = torch.device("cuda", 0)
torch.manual_seed(0)
ac = ActorCritic(O, O, A, actor_hidden_dims=[256, 256], critic_hidden_dims=[256, 256], activation="elu",
                 init_noise_std=1.0, spike_ts=5)
alg = PPO(ac, num_learning_epochs=5, num_mini_batches=4, clip_param=0.2, gamma=0.99, lam=0.95, value_loss_coef=1.0,
          entropy_coef=0.01, learning_rate=1e-3, max_grad_norm=1.0, use_clipped_value_loss=True, schedule="adaptive",
          desired_kl=0.01, device=dev, use_cuda_graph=(mode == "graph"))
alg.init_storage(N, S, [O], [None], [A])


def rollout():
    with torch.inference_mode():
        for s in range(S):
            obs = torch.randn(N, O, device=dev)
            alg.act(obs, obs)
            alg.process_env_step(torch.randn(N, device=dev), (torch.rand(N, device=dev) < 0.02).to(torch.uint8), {})
        alg.compute_returns(torch.randn(N, O, device=dev))


for _ in range(3):
    rollout()
    alg.update()
rollout()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    alg.update()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "mem" not in e.name.lower()[:6]]
ev.sort(key=lambda e: e.time_range.start)
t0, t1 = ev[0].time_range.start, max(e.time_range.end for e in ev)
by = defaultdict(lambda: [0, 0.0])
for e in ev:
    k = e.name.split("(")[0][:70]
    by[k][0] += 1
    by[k][1] += e.time_range.end - e.time_range.start
lines = [f"# PPO update timeline ({mode}): {len(ev)} device activities, span {(t1 - t0) / 1e3:.2f} ms "
         f"({(t1 - t0) / 20:.1f} us per minibatch step)", "",
         "| kernel | launches | total us | us/launch | us per minibatch step |", "|---|---:|---:|---:|---:|"]
for k, (n, us) in sorted(by.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"| `{k}` | {n} | {us:.0f} | {us / n:.1f} | {us / 20:.1f} |")
busy = sum(v[1] for v in by.values())
lines += ["", f"sum of activity durations {busy / 1e3:.2f} ms over a {(t1 - t0) / 1e3:.2f} ms span "
              f"(two streams overlap: actor chain + critic chain)"]
# one minibatch step in launch order (the 11th of 20), with start offsets
per = len(ev) // 20
seg = ev[10 * per: 11 * per] if per * 20 == len(ev) else ev[len(ev) // 2: len(ev) // 2 + 45]
base = seg[0].time_range.start
lines += ["", "## one minibatch step in start order (offset us, duration us, stream)", ""]
for e in seg:
    lines.append(f"    {e.time_range.start - base:8.1f}  {e.time_range.end - e.time_range.start:7.1f}  s{getattr(e, 'device_resource_id', '?')}  {e.name.split('(')[0][:80]}")
text = "\n".join(lines)
print(text)
if out:
    open(out, "w").write(text + "\n")
