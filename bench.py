#!/usr/bin/env python
"""bench.py — headline benchmark of the spiking actor-critic hot path.

BASELINE.json's metric has two halves; both are in the one JSON line rank 0 prints:

  (ii) PPO samples/s (forward + BPTT)  -> the line's `value` / `e2e` / `roofline` / `cpu_baseline`
       workload = BASELINE configs[1]: A1 AMP spiking actor-critic PPO update, 4096 envs x 24-step rollout.
       A "step" is one full PPO `update()` (5 epochs x 4 minibatches of 24 576 samples: actor forward with traces,
       critic, PPO losses, BPTT backward, grad-norm clip, Adam; reference rsl_rl/algorithms/ppo.py:120-187).
       N > 1: weak scaling, every rank owns its own 4096 environments, gradients all-reduced (`scaling: "weak"`).
  (i)  SNN policy env-steps/s (forward) -> the `fwd` block (value, e2e, roofline, cpu_baseline on configs[0]:
       the A1 spiking actor forward on 4096 observations; reference modules/actor_critic.py:300-320).

plus, at every N, a `strong` block: BASELINE configs[2] (MIT Humanoid, 4096 envs x 24 steps IN TOTAL, sharded
num_envs / world per GPU as SURVEY 8e partitions it) and `dp_check` (parameter checksums equal across ranks after
the timed updates).

  python bench.py --gpus 1 --steps 3 --warmup 3            # our arm
  python bench.py --impl reference --steps 3 --warmup 3    # CPU arm: the UNMODIFIED reference PPO.update (staged by
                                                           # oracle/make_ref.sh) or, if it is not staged, the oracle port
"""
from __future__ import annotations

import argparse
import glob
import json
import math
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

# BASELINE.json configs[1] (the headline; also the per-GPU workload of the weak-scaling curve)
WORKLOAD = {
    "name": "a1_amp_ppo_update_4096x24",
    "num_envs": 4096, "steps_per_env": 24, "obs_dim": 48, "act_dim": 12,
    "actor_hidden": [256, 256], "critic_hidden": [256, 256], "spike_ts": 5,
    "epochs": 5, "mini_batches": 4,
}
# BASELINE.json configs[0]: the forward half of the metric and its CPU-runnable case
FWD_WORKLOAD = {"name": "a1_rma_actor_fwd_4096", "batch": 4096, "obs_dim": 48, "act_dim": 12, "actor_hidden": [256, 256],
                "spike_ts": 5, "flops_per_sample": 2 * 5 * (480 * 256 + 256 * 256 + 256 * 120)}
# BASELINE.json configs[2]: the strong-scaling case, num_envs is the TOTAL over all ranks
STRONG_WORKLOAD = {
    "name": "mit_humanoid_ppo_update_4096x24_sharded",
    "num_envs": 4096, "steps_per_env": 24, "obs_dim": 38, "act_dim": 10,
    "actor_hidden": [256, 256, 256], "critic_hidden": [256, 256, 256], "spike_ts": 5,
    "epochs": 5, "mini_batches": 4,
}
METRIC = "ppo_samples_per_s_fwd_bptt"
UNIT = "samples/s"
CPU_SAMPLE_ROWS = 8192


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"bf16_tflops_sustained": p.get("bf16_tflops_sustained", 1400.0), "bf16_tflops": p.get("bf16_tflops", 1590.0),
                "hbm_gbs": p.get("hbm_gbs", 6650.0), "source": "MEASURED_PEAKS.json"}
    return {"bf16_tflops_sustained": 1400.0, "bf16_tflops": 1590.0, "hbm_gbs": 6650.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.path = f"/tmp/bench_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "20"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        with open(self.path) as f:
            for line in f:
                parts = [x.strip() for x in line.split(",")]
                if len(parts) < 8:
                    continue
                try:
                    sm.append(float(parts[1])); mx.append(float(parts[2]))
                except ValueError:
                    continue
                for n, v in zip(names, parts[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        try:
            os.remove(self.path)
        except OSError:
            pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------- CPU arms
def _bench_state_dict(W, seed=0):
    """Reference default-style init of the C2 actor-critic (same tensors for the reference modules and the port)."""
    from cases import Case, make_state_dict
    c = Case("bench", "amp", W["obs_dim"], tuple(W["actor_hidden"]), W["act_dim"], 1, spike_ts=W["spike_ts"], seed=seed)
    sd = make_state_dict(c)
    g = torch.Generator().manual_seed(seed)
    dims = [W["obs_dim"]] + W["critic_hidden"] + [1]
    for i in range(len(dims) - 1):
        b = 1.0 / math.sqrt(dims[i])
        sd[f"critic.{2 * i}.weight"] = (torch.rand(dims[i + 1], dims[i], generator=g) * 2 - 1) * b
        sd[f"critic.{2 * i}.bias"] = (torch.rand(dims[i + 1], generator=g) * 2 - 1) * b
    sd["std"] = torch.ones(W["act_dim"])
    return sd, g


def cpu_minibatch_runner(rows: int, seed: int = 0, device: str = "cpu"):
    """Oracle port of one PPO minibatch step (actor fwd with autograd graph, critic, losses, BPTT, clip, Adam)
    of the reference on the host cores (or, for `--impl port-gpu`, the same torch ops on the CUDA device: the stock
    PyTorch GPU path of SURVEY 8d).  Returns a callable running one step on `rows` samples."""
    from oracle import ppo_oracle as P
    W = WORKLOAD
    sd, g = _bench_state_dict(W, seed)
    sd = {k: v.to(device) for k, v in sd.items()}
    ac = P.OracleActorCritic(sd, spike_ts=W["spike_ts"])
    cfg = P.PPOConfig()
    opt = torch.optim.Adam(ac.parameters(), lr=cfg.learning_rate)
    A = W["act_dim"]
    obs = torch.randn(rows, W["obs_dim"], generator=g)
    batch = (obs, obs, torch.randn(rows, A, generator=g) * 0.5, torch.randn(rows, 1, generator=g),
             torch.randn(rows, 1, generator=g), torch.randn(rows, 1, generator=g),
             -0.5 * A * math.log(2 * math.pi) + torch.randn(rows, 1, generator=g) * 0.1 - 2.0,
             torch.randn(rows, A, generator=g) * 0.3, torch.ones(rows, A))
    batch = tuple(t.to(device) for t in batch)
    state = {"lr": cfg.learning_rate}

    def step():
        _, state["lr"] = P.ppo_minibatch_step(ac, opt, batch, cfg, state["lr"])

    return step


def reference_minibatch_runner(rows: int, seed: int = 0):
    """The UNMODIFIED reference on the host cores: AMP fork's ActorCritic + PPO + RolloutStorage (staged under
    oracle/_ref by oracle/make_ref.sh; AMP-rl/algorithms/ppo.py:120-187), one `PPO.update()` with one epoch and one
    minibatch of `rows` samples = one minibatch step (incl. the reference's own randperm + nine gathers).
    Returns None when the reference is not staged."""
    from oracle import ref_loader as R
    if not R.available("amp"):
        return None
    W = WORKLOAD
    mods, algs, _ = R.load_all("amp")
    sd, g = _bench_state_dict(W, seed)
    with R.quiet():
        ac = mods.ActorCritic(W["obs_dim"], W["obs_dim"], W["act_dim"], actor_hidden_dims=list(W["actor_hidden"]),
                              critic_hidden_dims=list(W["critic_hidden"]), activation="elu", init_noise_std=1.0)
    ac.actor.encoder.device = ac.actor.snn.device = "cpu"        # the ctor hard-codes "cuda" for its zeros
    ac.load_state_dict(sd)
    alg = algs.PPO(ac, num_learning_epochs=1, num_mini_batches=1, clip_param=0.2, gamma=0.99, lam=0.95,
                   value_loss_coef=1.0, entropy_coef=0.01, learning_rate=1e-3, max_grad_norm=1.0,
                   use_clipped_value_loss=True, schedule="adaptive", desired_kl=0.01, device="cpu")
    S = 8
    N = rows // S
    A = W["act_dim"]
    alg.init_storage(N, S, [W["obs_dim"]], [None], [A])
    st = alg.storage
    st.observations.copy_(torch.randn(S, N, W["obs_dim"], generator=g))
    st.actions.copy_(torch.randn(S, N, A, generator=g) * 0.5)
    st.values.copy_(torch.randn(S, N, 1, generator=g))
    st.returns.copy_(torch.randn(S, N, 1, generator=g))
    st.advantages.copy_(torch.randn(S, N, 1, generator=g))
    st.actions_log_prob.copy_(-0.5 * A * math.log(2 * math.pi) + torch.randn(S, N, 1, generator=g) * 0.1 - 2.0)
    st.mu.copy_(torch.randn(S, N, A, generator=g) * 0.3)
    st.sigma.fill_(1.0)

    def step():
        with R.quiet():                      # the AMP fork's forward prints on every call
            alg.update()

    return step


def time_cpu(rows: int, steps: int, warmup: int):
    """(samples/s, s/step, kind) of one PPO minibatch step on `rows` rows on all host cores."""
    torch.set_num_threads(os.cpu_count() or 1)
    step = reference_minibatch_runner(rows)
    kind = "reference"
    if step is None:
        step, kind = cpu_minibatch_runner(rows), "port"
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return rows / dt, dt, kind


def time_cpu_forward(batch: int, reps: int = 5, warmup: int = 2):
    """configs[0]: the reference's PopSpikeActor forward (obs -> mu, std) on `batch` observations, host cores, under
    inference_mode, stdout silenced; best of `reps`.  Falls back to the oracle port when not staged."""
    from cases import Case, make_state_dict
    from oracle import ref_loader as R
    W = FWD_WORKLOAD
    torch.set_num_threads(os.cpu_count() or 1)
    c = Case("bench_fwd", "rma", W["obs_dim"], tuple(W["actor_hidden"]), W["act_dim"], 1, spike_ts=W["spike_ts"], seed=0)
    sd = {k[len("actor."):]: v for k, v in make_state_dict(c).items() if k.startswith("actor.")}
    obs = torch.randn(batch, W["obs_dim"], generator=torch.Generator().manual_seed(0))
    if R.available("rma"):
        mods = R.load("rma", "modules.actor_critic")
        with R.quiet():
            actor = mods.PopSpikeActor(W["obs_dim"], W["act_dim"], 10, 10, list(W["actor_hidden"]), (-5, 5),
                                       math.sqrt(0.15), spike_ts=W["spike_ts"], device="cpu")
        actor.load_state_dict(sd)
        kind = "reference"

        def fwd():
            with R.quiet(), torch.inference_mode():
                actor(obs)
    else:
        from helpers import oracle_params
        from oracle import snn_oracle as O
        prm, kind = oracle_params(c), "port"

        def fwd():
            with torch.inference_mode():
                O.actor_forward(obs, prm)
    for _ in range(warmup):
        fwd()
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        fwd()
        best = min(best, time.perf_counter() - t0)
    return batch / best, best, kind


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    value, dt, kind = time_cpu(CPU_SAMPLE_ROWS, args.steps, args.warmup)
    cores = torch.get_num_threads()
    what = ("the UNMODIFIED reference (AMP fork ActorCritic + PPO.update + RolloutStorage, staged by oracle/make_ref.sh)"
            if kind == "reference" else "oracle port of the reference (torch CPU fp32 autograd)")
    sample = (f"one PPO minibatch step (actor fwd+BPTT, critic, losses, clip, Adam) on {CPU_SAMPLE_ROWS} of the "
              f"24576 minibatch rows per timed step; {what}")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD["name"], **{k: v for k, v in WORKLOAD.items() if k != "name"}},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if not args.quick:
        fv, fdt, fkind = time_cpu_forward(FWD_WORKLOAD["batch"])
        line["fwd"] = {"metric": "snn_policy_env_steps_per_s_fwd", "value": fv, "unit": "env-steps/s",
                       "ms_per_call": fdt * 1e3, "kind": fkind, "cores": cores, "workload": FWD_WORKLOAD["name"]}
    print(json.dumps(line))
    return 0


def run_port_gpu(args):
    """The oracle port (the reference's torch ops, op for op) on the CUDA device with stock PyTorch kernels: the
    same-device comparison SURVEY 8d asks for next to the CPU arm.  Full 24 576-row minibatch steps, CUDA events.
    A baseline leg like `--impl reference`: none of this repo's kernels run here."""
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    if not torch.cuda.is_available():
        raise RuntimeError("--impl port-gpu needs a CUDA device")
    torch.backends.cuda.matmul.allow_tf32 = False       # fp32 like the CPU arm and like our parity mode
    torch.backends.cudnn.allow_tf32 = False
    rows = WORKLOAD["num_envs"] * WORKLOAD["steps_per_env"] // WORKLOAD["mini_batches"]
    step = cpu_minibatch_runner(rows, device="cuda")
    for _ in range(max(args.warmup, 3)):
        step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    line = {
        "impl": "port-gpu", "metric": METRIC, "value": rows / (ms * 1e-3), "unit": UNIT, "n_gpus": 1,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD["name"], "rows_per_step": rows,
                   "note": "one PPO minibatch step (actor fwd + BPTT by autograd, critic, losses, clip, Adam) of the "
                           "oracle port with stock PyTorch CUDA kernels, TF32 off; a step here is ONE minibatch, "
                           "our arm's step is a whole update of 20"},
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------- our arm
def latest_traffic():
    """The newest committed ncu capture summary (profiles/r*_traffic.json): DRAM bytes / tensor-pipe % per kernel, with
    the commit it was captured at.  ncu numbers cannot be produced inside a bench run (never time under a profiler)."""
    paths = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")))
    if not paths:
        return None, None
    with open(paths[-1]) as f:
        return json.load(f), os.path.relpath(paths[-1], ROOT)


def run_ours(args):
    import ctypes as C

    import torch.distributed as dist

    from fullysnnforleggedrobots_b200 import PPO, ActorCritic, ActorCriticHumanoid, _lib
    from fullysnnforleggedrobots_b200 import dist as snn_dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the spiking actor has no CPU fallback")
    L = _lib.lib()
    if L.snn_build_flags() != 0 or os.environ.get("SNN_DBG"):
        raise RuntimeError("bench.py refuses to time a tools build (-DSNN_ABLATE) or a run with SNN_DBG set: "
                           f"build_flags={L.snn_build_flags()} SNN_DBG={os.environ.get('SNN_DBG')!r}")
    if os.environ.get("SNN_GRAPH_FALLBACK", "0") == "1":
        raise RuntimeError("bench.py refuses SNN_GRAPH_FALLBACK=1: a failed CUDA-graph capture must be an error here")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    reducer = snn_dist.init_from_env("nccl") if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def make_job(W, envs_local, cls, seed_off):
        """ActorCritic + PPO + a filled rollout of `envs_local` environments for workload W (host copies pinned)."""
        torch.manual_seed(0)
        ac = cls(W["obs_dim"], W["obs_dim"], W["act_dim"], actor_hidden_dims=W["actor_hidden"],
                 critic_hidden_dims=W["critic_hidden"], activation="elu", init_noise_std=1.0, spike_ts=W["spike_ts"])
        alg = PPO(ac, num_learning_epochs=W["epochs"], num_mini_batches=W["mini_batches"], clip_param=0.2, gamma=0.99,
                  lam=0.95, value_loss_coef=1.0, entropy_coef=0.01, learning_rate=1e-3, max_grad_norm=1.0,
                  use_clipped_value_loss=True, schedule="adaptive", desired_kl=0.01, device=dev, data_parallel=reducer)
        N, S, O, A = envs_local, W["steps_per_env"], W["obs_dim"], W["act_dim"]
        alg.init_storage(N, S, [O], [None], [A])
        g = torch.Generator().manual_seed(1234 + seed_off + rank)
        host = {
            "obs": torch.randn(S, N, O, generator=g).pin_memory(),
            "rewards": torch.randn(S, N, generator=g).pin_memory(),
            "dones": (torch.rand(S, N, generator=g) < 0.02).to(torch.uint8).pin_memory(),
            "last_obs": torch.randn(N, O, generator=g).pin_memory(),
        }
        with torch.inference_mode():
            for s in range(S):
                obs = host["obs"][s].to(dev, non_blocking=True)
                alg.act(obs, obs)
                alg.process_env_step(host["rewards"][s].to(dev), host["dones"][s].to(dev), {})
            alg.compute_returns(host["last_obs"].to(dev))
        for name in ("actions", "values", "actions_log_prob", "mu", "sigma"):
            host[name] = getattr(alg.storage, name).detach().cpu().pin_memory()
        torch.cuda.synchronize()
        return ac, alg, host

    W = WORKLOAD
    N, S, O, A = W["num_envs"], W["steps_per_env"], W["obs_dim"], W["act_dim"]
    ac, alg, host = make_job(W, N, ActorCritic, 0)
    st = alg.storage
    samples_per_step = N * S * W["epochs"]          # per rank

    # ---- device-resident leg
    def step_resident():
        alg.update()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()          # started before warm-up: nvidia-smi needs ~0.5 s before its first sample
    for _ in range(args.warmup):
        step_resident()
    launches0 = L.snn_launch_count() + alg.replayed_kernel_launches
    ms_total = timed(step_resident, args.steps)
    launches = L.snn_launch_count() + alg.replayed_kernel_launches - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = ms_total / args.steps
    value = world * samples_per_step / (ms_per_step * 1e-3)
    graphs_used = bool(alg.use_cuda_graph and len(alg._graphs) > 0)
    if not graphs_used:
        raise RuntimeError("the timed update did not replay CUDA graphs (capture failed or was disabled)")

    # ---- dp_check: after the timed updates every rank must hold bit-identical parameters (the gradients it applied
    #      came out of one all-reduce; the rollouts differ per rank, so equality is not trivial)
    flat = alg.optimizer.flat_params
    chk = torch.stack([flat.double().sum(), flat.view(torch.int32).to(torch.int64).sum().double()])
    if world > 1:
        allc = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk)
        dp_check = {"ranks": world, "params_bit_identical": bool(all(torch.equal(c, allc[0]) for c in allc)),
                    "param_sum": float(chk[0].item()), "dp_graph": alg.dp_graph_mode, "dp_graph_note": alg.dp_graph_note,
                    "dp_collective": "peer-memory sum inside the optimizer kernel (snn_clip_adam_kl_p2p)" if alg.dp_p2p
                    else "nccl all-reduce (" + getattr(alg.dp, "p2p_note", "") + ")"}
    else:
        dp_check = {"ranks": 1, "params_bit_identical": True, "param_sum": float(chk[0].item()), "dp_graph": "n/a",
                    "dp_graph_note": ""}

    # ---- per-kernel device times for the roofline: the SAME steps, launched eagerly (events cannot be read
    #      back from inside a replayed CUDA graph); kernel durations do not depend on how they were launched
    graph_flag = alg.use_cuda_graph
    alg.use_cuda_graph = False
    step_resident()
    L.snn_profile_enable(1)
    ms_prof_total = timed(step_resident, args.steps)
    pms32, pfl32, pln32 = (C.c_double * 32)(), (C.c_double * 32)(), (C.c_int64 * 32)()
    L.snn_profile_read(pms32, pfl32, pln32)
    # layers 0..3 = the spiking actor, 4..7 = the dense critic (same kernel family, MLP epilogues)
    pms = [sum(pms32[k + 4 * l] for l in range(4)) for k in range(4)]
    pfl = [sum(pfl32[k + 4 * l] for l in range(4)) for k in range(4)]
    pln = [sum(pln32[k + 4 * l] for l in range(4)) for k in range(4)]
    cms = [sum(pms32[k + 4 * l] for l in range(4, 8)) for k in range(4)]
    cfl = [sum(pfl32[k + 4 * l] for l in range(4, 8)) for k in range(4)]
    cln = [sum(pln32[k + 4 * l] for l in range(4, 8)) for k in range(4)]
    per_layer = {f"{['fwd', 'dx', 'dw'][k]}_layer{l}": {"us_per_launch": 1e3 * pms32[k + 4 * l] / pln32[k + 4 * l],
                                                        "algo_tflops": pfl32[k + 4 * l] / (pms32[k + 4 * l] * 1e-3) / 1e12}
                 for l in range(8) for k in range(3) if pln32[k + 4 * l] > 0}
    L.snn_profile_enable(0)
    alg.use_cuda_graph = graph_flag

    # ---- end-to-end leg: host buffers in, loss out, every step
    h2d_fields = ["obs", "rewards", "dones", "actions", "values", "actions_log_prob", "mu", "sigma", "last_obs"]
    h2d_bytes = sum(host[k].numel() * host[k].element_size() for k in h2d_fields)

    def step_e2e():
        st.observations.copy_(host["obs"], non_blocking=True)
        st.rewards.copy_(host["rewards"].unsqueeze(-1), non_blocking=True)
        st.dones.copy_(host["dones"].unsqueeze(-1), non_blocking=True)
        st.actions.copy_(host["actions"], non_blocking=True)
        st.values.copy_(host["values"], non_blocking=True)
        st.actions_log_prob.copy_(host["actions_log_prob"], non_blocking=True)
        st.mu.copy_(host["mu"], non_blocking=True)
        st.sigma.copy_(host["sigma"], non_blocking=True)
        last = host["last_obs"].to(dev, non_blocking=True)
        with torch.inference_mode():
            alg.compute_returns(last)
        return alg.update()                      # returns two host floats (device -> host read of the losses)

    if args.quick:
        ms_e2e, e2e_value = float("nan"), float("nan")
    else:
        step_e2e()
        ms_e2e = timed(step_e2e, args.steps) / args.steps
        e2e_value = world * samples_per_step / (ms_e2e * 1e-3)

    # ---- forward half of the metric: the spiking actor's rollout forward at B = 4096 (configs[0] dims == C2's actor)
    FW = FWD_WORKLOAD
    Bf = FW["batch"]
    nrep = 2 if args.quick else 50
    obs_dev = host["obs"][0][:Bf].to(dev)
    obs_pin = host["obs"][0][:Bf].contiguous().pin_memory()
    mu_pin = torch.empty(Bf, A).pin_memory()

    def step_act():                                  # obs resident in HBM -> mu, std resident
        with torch.inference_mode():
            ac.act_inference(obs_dev)

    def step_act_e2e():                              # host obs in, host mu out, through the module's public call
        with torch.inference_mode():
            mu_pin.copy_(ac.act_inference(obs_pin.to(dev, non_blocking=True)), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(2 if args.quick else 10):
        step_act()
    fl0 = L.snn_launch_count()
    ms_act = timed(step_act, nrep) / nrep
    fwd_launches = (L.snn_launch_count() - fl0) // nrep
    step_act_e2e()
    ms_act_e2e = timed(step_act_e2e, nrep) / nrep
    ms_act_eager = ms_act
    # the same call replayed as a CUDA graph (how a rollout loop runs it: no per-call host work; PPO.act does the same)
    with torch.inference_mode():
        gside = torch.cuda.Stream()
        gside.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(gside):
            ac.act_inference(obs_dev)
        torch.cuda.current_stream().wait_stream(gside)
        torch.cuda.synchronize()
        fwd_graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(fwd_graph):
            mu_static = ac.act_inference(obs_dev)
    for _ in range(3):
        fwd_graph.replay()
    ms_act_graph = timed(fwd_graph.replay, nrep) / nrep
    with torch.inference_mode():
        if not torch.equal(mu_static, ac.act_inference(obs_dev)):
            raise RuntimeError("graph replay of act_inference differs from the eager call")
    ms_act = min(ms_act_eager, ms_act_graph)

    # full rollout step PPO.act (actor forward + sample + log-prob + critic), replayed as one CUDA graph
    def step_ppo_act():
        with torch.inference_mode():
            alg.act(obs_dev, obs_dev)
            alg.transition.clear()

    for _ in range(2 if args.quick else 6):
        step_ppo_act()
    ms_pact = timed(step_ppo_act, nrep) / nrep

    # ---- strong scaling: BASELINE configs[2] / SURVEY C3, 4096 environments IN TOTAL, num_envs / world per rank
    SW = STRONG_WORKLOAD
    strong = None
    if not args.quick:
        lo, hi = snn_dist.shard_range(SW["num_envs"], rank, world)
        if (hi - lo) * SW["steps_per_env"] % SW["mini_batches"] != 0 or SW["num_envs"] % world != 0:
            raise RuntimeError("strong-scaling workload does not shard evenly over this world size")
        _, salg, _ = make_job(SW, hi - lo, ActorCriticHumanoid, 77)
        for _ in range(max(args.warmup, 3)):
            salg.update()
        ms_s = timed(salg.update, args.steps) / args.steps
        strong = {"workload": SW["name"], "scaling": "strong", "value": SW["num_envs"] * SW["steps_per_env"] * SW["epochs"] / (ms_s * 1e-3),
                  "unit": UNIT, "ms_per_step": ms_s, "envs_total": SW["num_envs"], "envs_per_gpu": hi - lo,
                  "minibatch_rows_per_gpu": (hi - lo) * SW["steps_per_env"] // SW["mini_batches"],
                  "flops_per_sample_fwd_bptt": 6645760,
                  "note": "total work fixed: 98304 samples x 5 epochs per update whatever N; per-rank minibatches shrink to "
                          "24576 / N rows, so every step is 20 x (~37 launch-bound kernels + one gradient all-reduce)"}

    def leave():
        """Multi-rank exit.  The step graphs hold captured NCCL kernels; tearing the communicator down under them
        (destroy_process_group) can block forever (seen at N = 2: the JSON line was out, the processes never left).
        All ranks meet once more, drain the device and leave without the NCCL teardown."""
        if world > 1:
            sys.stdout.flush()
            sys.stderr.flush()
            dist.barrier()
            torch.cuda.synchronize()
            os._exit(0)
        return 0

    if rank != 0:
        return leave()

    pk = peaks()
    cats = ["fwd_nm_gemm", "dx_nm_gemm", "dw_gemm", "other_simt"]
    per_cat = {c: {"ms": pms[i], "launches": int(pln[i]), "algo_tflops": (pfl[i] / (pms[i] * 1e-3) / 1e12) if pms[i] > 0 else 0.0}
               for i, c in enumerate(cats)}
    # dominant kernel = the single (kind, layer) launch site of the actor with the largest total device time
    sites = [(k, l) for l in range(4) for k in range(3) if pln32[k + 4 * l] > 0]
    dk, dl = max(sites, key=lambda kl: pms32[kl[0] + 4 * kl[1]])
    di = dk + 4 * dl
    dom_name = f"{['fwd', 'dx', 'dw'][dk]}_layer{dl}"
    dom_kernel = (["nm_gemm_kernel<FWD>", "nm_gemm_kernel<DX>" if dl > 0 else "nm_gemm_kernel<DX_ENC>"][dk] + f" (actor layer {dl})"
                  if dk < 2 else "dw_gemm_multi_kernel (weight gradients of all actor layers, one launch)")
    achieved = pfl32[di] / (pms32[di] * 1e-3) / 1e12
    traffic, pipe_pct, smem_pct = None, None, None
    tj, tpath = latest_traffic()
    hbm_view = {}
    if tj is not None:
        tk = tj["kernels"].get(dom_name)
        if tk:
            traffic, pipe_pct = tk["dram_bytes_per_launch"], tk["tensor_pipe_active_pct"]
            smem_pct = tk.get("smem_pipe_tensor_pct", 0.0) + tk.get("smem_pipe_lsu_pct", 0.0)
        # secondary bound: HBM.  ncu DRAM bytes per launch (committed capture) over this run's CUDA-event kernel times.
        for name, v in per_layer.items():
            tkk = tj["kernels"].get(name)
            if tkk and v["us_per_launch"] > 0:
                gbs = tkk["dram_bytes_per_launch"] / (v["us_per_launch"] * 1e-6) / 1e9
                hbm_view[name] = {"dram_gbs": gbs, "frac_of_hbm_peak": gbs / pk["hbm_gbs"]}
    tc_ms = pms[0] + pms[1] + pms[2]
    tc_fl = pfl[0] + pfl[1] + pfl[2]
    roofline = {
        "bound": "tensor", "kernel": dom_kernel, "achieved": achieved, "peak": pk["bf16_tflops_sustained"],
        "unit": "TFLOP/s", "frac": achieved / pk["bf16_tflops_sustained"], "traffic": traffic,
        "us_per_launch": 1e3 * pms32[di] / pln32[di], "algo_flops_per_launch": pfl32[di] / pln32[di],
        "ncu_tensor_pipe_active_pct": pipe_pct, "ncu_smem_data_pipe_pct": smem_pct,
        "ncu_capture": {"file": tpath, "commit": tj.get("commit") if tj else None,
                        "note": "traffic / tensor-pipe % / shared-memory data-pipe % are read from this committed ncu "
                                "capture (a profiler never runs inside a bench run); achieved / frac are measured live"},
        "peak_source": pk["source"] + " (bf16_tflops_sustained: kernel timed inside a long step)",
        "note": "achieved = ALGORITHMIC flops (2*B*K*N*T per layer GEMM, one product per MAC) / CUDA-event kernel time; "
                "fp32-parity mode issues 3 bf16-plane MMAs per algorithmic MAC in the forward / dX kernels and 2 in dW (g_c "
                "hi + lo planes x exact spikes), so issued tensor work is 3x / 2x 'achieved'; "
                "kernel times from an eager re-run of the timed steps (the timed region itself replays CUDA graphs)",
        "whole_step": {"algo_tflops": samples_per_step * 5345280 / (ms_per_step * 1e-3) / 1e12,
                       "frac": samples_per_step * 5345280 / (ms_per_step * 1e-3) / 1e12 / pk["bf16_tflops_sustained"],
                       "note": "per GPU: actor fwd+BPTT flops (5 345 280 per sample) over the whole timed update"},
        "all_tensor_kernels": {"ms": tc_ms, "algo_tflops": tc_fl / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.0,
                               "share_of_step": tc_ms / ms_total,
                               "eager_profiled_ms_per_step": ms_prof_total / args.steps},
        "hbm_view": {"peak_gbs": pk["hbm_gbs"], "per_kernel": hbm_view},
        "per_kernel": per_cat,
        "per_layer": per_layer,
        "critic_tensor_kernels": {"ms": sum(cms[:3]), "launches": int(sum(cln[:3])),
                                  "algo_tflops": sum(cfl[:3]) / (sum(cms[:3]) * 1e-3) / 1e12 if sum(cms[:3]) > 0 else 0.0},
    }
    fwd_tf = Bf * FW["flops_per_sample"] / (ms_act * 1e-3) / 1e12
    fwd = {
        "metric": "snn_policy_env_steps_per_s_fwd", "workload": FW["name"], "unit": "env-steps/s",
        "value": world * Bf / (ms_act * 1e-3), "ms_per_call": ms_act, "batch_per_gpu": Bf, "gpu_launches_per_call": int(fwd_launches),
        "ms_per_call_eager": ms_act_eager, "ms_per_call_graph_replay": ms_act_graph,
        "e2e": {"value": world * Bf / (ms_act_e2e * 1e-3), "unit": "env-steps/s", "ms_per_call": ms_act_e2e,
                "h2d_bytes_per_step": Bf * O * 4, "d2h_bytes_per_step": Bf * A * 4},
        "roofline": {"bound": "tensor", "achieved": fwd_tf, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                     "frac": fwd_tf / pk["bf16_tflops_sustained"], "traffic": None,
                     "algo_flops_per_call": Bf * FW["flops_per_sample"],
                     "note": "per GPU; the whole forward call (one launch: snn_fwd_act), 2*T*sum K*N flops per sample; value = the "
                             "faster of the eager module call and its CUDA-graph replay (both reported); at 4096 environments "
                             "one CTA per 32-sample tile streams all weight planes from L2 (9 GFLOP = 6.4 us at peak)"},
        "ppo_act": {"value": world * N / (ms_pact * 1e-3), "ms_per_call": ms_pact,
                    "note": "the whole PPO.act rollout step (actor forward, sample, log-prob, critic value), one CUDA-graph replay"},
    }
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": W["name"], **{k: v for k, v in W.items() if k != "name"},
                   "arithmetic": "fp32-parity: weights as three exact bf16 planes (hi+mid+lo == W), spikes exact, fp32 "
                                 "accumulation on the tensor cores; backward operands as hi/lo bf16 pairs (3 products)",
                   "parallelism": f"dp{world}", "envs_per_gpu": N,
                   "l2": "inputs larger than L2: 54 MB rollout + 0.17 GB traces + 0.7 GB gradient planes per minibatch vs 126 MB L2",
                   "cuda_graphs": graphs_used, "build_flags": int(L.snn_build_flags())},
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": 8},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "fwd": fwd,
        "strong": strong,
        "dp_check": dp_check,
    }
    if world == 1 and not args.quick:
        cpu_value, cpu_dt, kind = time_cpu(CPU_SAMPLE_ROWS, 3, 2)
        what = ("UNMODIFIED reference PPO.update (AMP fork, staged by oracle/make_ref.sh)" if kind == "reference"
                else "oracle port (torch CPU fp32 autograd)")
        line["cpu_baseline"] = {
            "value": cpu_value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
            "sample": f"one PPO minibatch step on {CPU_SAMPLE_ROWS} of 24576 rows, 2 warm-up + 3 timed, {what}; "
                      f"{cpu_dt:.2f} s/step"}
        fv, fdt, fkind = time_cpu_forward(Bf)
        fwd["cpu_baseline"] = {"value": fv, "unit": "env-steps/s", "cores": torch.get_num_threads(), "kind": fkind,
                               "sample": f"PopSpikeActor forward on the same {Bf} observations (configs[0]), inference_mode, "
                                         f"best of 5; {fdt * 1e3:.1f} ms per call"}
    print(json.dumps(line))
    return leave()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "port-gpu"])
    ap.add_argument("--quick", action="store_true", help="resident leg only (for ncu runs): no e2e / strong / CPU legs")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.impl == "port-gpu":
        return run_port_gpu(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
