"""Time the rollout step PPO.act (one CUDA-graph replay) with CUDA events and check its critic branch against torch.
A/B of the critic branch: SNN_ACT_CRITIC=torch python tools/time_act.py  vs  SNN_ACT_CRITIC=tc (default)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from fullysnnforleggedrobots_b200 import PPO, ActorCritic

N, O, A = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, 48, 12
REP = 300
dev = torch.device("cuda", 0)
torch.manual_seed(0)
ac = ActorCritic(O, O, A, actor_hidden_dims=[256, 256], critic_hidden_dims=[256, 256], activation="elu",
                 init_noise_std=1.0, spike_ts=5)
alg = PPO(ac, num_learning_epochs=5, num_mini_batches=4, device=dev)
alg.init_storage(N, 24, [O], [None], [A])
obs = torch.randn(N, O, device=dev)
with torch.inference_mode():
    for _ in range(10):
        alg.act(obs, obs)
        v = alg.transition.values.clone()
        alg.transition.clear()
    ref = ac.evaluate(obs)
    err = float((v - ref).abs().max() / ref.abs().max())
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(REP):
            alg.act(obs, obs)
            alg.transition.clear()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / REP)
    # the replayed graph alone (no input copy, no output clone, no host logic of PPO.act)
    g = alg._act_state["graph"]
    gbest = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(REP):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        gbest = min(gbest, e0.elapsed_time(e1) / REP)
print(json.dumps({"graph_only_us": gbest * 1e3, "critic_branch": "tc" if alg._act_state.get("ce") is not None else "torch", "envs": N,
                  "ppo_act_us": best * 1e3, "env_steps_per_s": N / (best * 1e-3), "value_rel_err_vs_torch": err}))
