"""Kernel timeline of one rollout step (PPO.act at 4096 envs, CUDA-graph replay) from CUPTI."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import ProfilerActivity, profile

from fullysnnforleggedrobots_b200 import PPO, ActorCritic

N, O, A = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, 48, 12
dev = torch.device("cuda", 0)
torch.manual_seed(0)
ac = ActorCritic(O, O, A, actor_hidden_dims=[256, 256], critic_hidden_dims=[256, 256], activation="elu",
                 init_noise_std=1.0, spike_ts=5)
alg = PPO(ac, num_learning_epochs=5, num_mini_batches=4, device=dev)
alg.init_storage(N, 24, [O], [None], [A])
obs = torch.randn(N, O, device=dev)
with torch.inference_mode():
    for _ in range(8):
        alg.act(obs, obs)
        alg.transition.clear()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            alg.act(obs, obs)
            alg.transition.clear()
        torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
per = len(ev) // 3
seg = ev[per:2 * per]
base = seg[0].time_range.start
print(f"{per} device activities per act(); span {(seg[-1].time_range.end - base):.1f} us")
for e in seg:
    print(f"  {e.time_range.start - base:8.1f} {e.time_range.end - e.time_range.start:7.1f}  {e.name.split('(')[0][:90]}")
