"""Wall/device time of AMPPPO.update() at the C2 size (4096 envs x 24 steps, 5 x 4 minibatches) with the AMP fork's
discriminator dims (input 2 x 30, hidden [1024, 512]; a1_amp_config.py) next to PPO.update() on the same rollout."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from fullysnnforleggedrobots_b200 import AMPPPO, PPO, ActorCritic
from fullysnnforleggedrobots_b200.amp import AMPDiscriminator, Normalizer

N, S, O, A, D = 4096, 24, 48, 12, 30
dev = torch.device("cuda", 0)


class Expert:
    def __init__(self):
        self.s = torch.randn(200000, D, device=dev)
        self.n = torch.randn(200000, D, device=dev)

    def feed_forward_generator(self, num, mbs):
        for _ in range(num):
            idx = torch.randint(0, self.s.shape[0], (mbs,), device=dev)
            yield self.s[idx], self.n[idx]


def make(kind):
    torch.manual_seed(0)
    ac = ActorCritic(O, O, A, actor_hidden_dims=[256, 256], critic_hidden_dims=[256, 256], activation="elu",
                     init_noise_std=1.0, spike_ts=5)
    kw = dict(num_learning_epochs=5, num_mini_batches=4, clip_param=0.2, gamma=0.99, lam=0.95, value_loss_coef=1.0,
              entropy_coef=0.01, learning_rate=1e-3, max_grad_norm=1.0, use_clipped_value_loss=True, schedule="adaptive",
              desired_kl=0.01, device=dev)
    if kind == "ppo":
        alg = PPO(ac, **kw)
    else:
        disc = AMPDiscriminator(2 * D, 2.0, [1024, 512], dev)
        alg = AMPPPO(ac, disc, Expert(), Normalizer(D), amp_replay_buffer_size=1000000, min_std=None, **kw)
    alg.init_storage(N, S, [O], [None], [A])
    return alg


def rollout(alg, amp):
    with torch.inference_mode():
        for s in range(S):
            obs = torch.randn(N, O, device=dev)
            if amp:
                alg.act(obs, obs, torch.randn(N, D, device=dev))
                alg.process_env_step(torch.randn(N, device=dev), (torch.rand(N, device=dev) < 0.02).to(torch.uint8), {},
                                     torch.randn(N, D, device=dev))
            else:
                alg.act(obs, obs)
                alg.process_env_step(torch.randn(N, device=dev), (torch.rand(N, device=dev) < 0.02).to(torch.uint8), {})
        alg.compute_returns(torch.randn(N, O, device=dev))


for kind in ("ppo", "amp"):
    alg = make(kind)
    for _ in range(3):
        rollout(alg, kind == "amp")
        alg.update()
    ts = []
    for _ in range(3):
        rollout(alg, kind == "amp")
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = alg.update()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(kind, "update ms:", [round(t, 2) for t in ts], "returns", [round(float(x), 4) for x in out])

if "--profile" in sys.argv:
    from collections import defaultdict
    from torch.profiler import ProfilerActivity, profile
    rollout(alg, True)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        alg.update()
        torch.cuda.synchronize()
    by = defaultdict(lambda: [0, 0.0])
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            k = e.name.split("(")[0][:80]
            by[k][0] += 1
            by[k][1] += e.time_range.end - e.time_range.start
    for k, (n, us) in sorted(by.items(), key=lambda kv: -kv[1][1])[:14]:
        print(f"{us / 1e3:9.2f} ms {n:6d}  {k}")
