"""Throughput of the spiking actor on every BASELINE.json configuration (C1..C5 + the forks' real dims).

For each (dims, T, B): forward-only env-steps/s (`PopSpikeActor.mean_only` under inference_mode = what `act`
runs) and forward + BPTT samples/s (training forward with traces, backward of sum(mu*G)), CUDA events, through the
public module.  Algorithmic TFLOP/s uses SURVEY §8d's count (2*T*sum K*N forward; backward = dW + dX, layer-0 dX only
when the actor input needs a gradient).  Writes a markdown table to stdout / argv[1].

    python tools/sweep_configs.py [out.md] [--quick]
"""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from fullysnnforleggedrobots_b200 import PopSpikeActor

QUICK = "--quick" in sys.argv
OUT = next((a for a in sys.argv[1:] if not a.startswith("--")), None)
MEM_CAP = 60e9          # skip a point whose traces would not fit comfortably (bounded memory sweep)


def flops(O, H, A, T, dobs):
    dims = [O * 10] + list(H) + [A * 10]
    macs = [dims[i] * dims[i + 1] for i in range(len(dims) - 1)]
    fwd = 2 * T * sum(macs)
    bwd = fwd + 2 * T * (sum(macs[1:]) + (macs[0] if dobs else 0))
    return fwd, fwd + bwd


def timed(fn, reps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def run(name, O, H, A, T, B_fwd, B_bwd, variant=None, dobs=False):
    kw = {"enc_eps": 1e-5} if variant == "humanoid" else ({"dec_tanh": True} if variant == "cassie" else {})
    torch.manual_seed(0)
    actor = PopSpikeActor(O, A, 10, 10, list(H), (-5, 5), math.sqrt(0.15), spike_ts=T, **kw).cuda()
    f_fwd, f_all = flops(O, H, A, T, dobs)
    rows = []
    for B in B_fwd:
        obs = torch.randn(B, O, device="cuda")

        def fwd():
            with torch.inference_mode():
                actor.mean_only(obs)
        try:
            for _ in range(3):
                fwd()
            ms = timed(fwd, 5 if B >= 262144 else 20)
            rows.append((name, T, B, "fwd (act)", ms, B / ms * 1e3, B * f_fwd / ms / 1e9))
        except Exception as e:  # noqa: BLE001
            rows.append((name, T, B, "fwd (act)", None, None, str(e)[:60]))
        del obs
    width = sum(((h + 127) // 128) * 128 for h in list(H) + [A * 10])
    for B in B_bwd:
        need = B * T * width * (2 + 4 + 4) * 1.1      # 16-bit voltage codes + two bf16 g_c planes (all layers kept for dW)
        if need > MEM_CAP:
            rows.append((name, T, B, "fwd+BPTT", None, None, f"skipped: ~{need / 1e9:.0f} GB of traces"))
            continue
        obs = torch.randn(B, O, device="cuda", requires_grad=dobs)
        G = torch.randn(B, A, device="cuda")

        def step():
            for p in actor.parameters():
                p.grad = None
            obs.grad = None
            mu, _ = actor(obs)
            (mu * G).sum().backward()
        try:
            for _ in range(2):
                step()
            ms = timed(step, 3 if B >= 131072 else 8)
            rows.append((name, T, B, "fwd+BPTT" + (" (+d obs)" if dobs else ""), ms, B / ms * 1e3, B * f_all / ms / 1e9))
        except Exception as e:  # noqa: BLE001
            rows.append((name, T, B, "fwd+BPTT", None, None, str(e)[:60]))
        del obs, G
        torch.cuda.empty_cache()
    del actor
    torch.cuda.empty_cache()
    return rows


def main():
    rows = []
    rows += run("C1 (48 -> 256-256 -> 12)", 48, [256, 256], 12, 5, [4096], [24576], "amp")
    rows += run("AMP actual (42 -> 512-256-128 -> 12; 5480 envs x 36)", 42, [512, 256, 128], 12, 5, [5480], [49320], "amp")
    rows += run("C3 Humanoid (38 -> 256-256-256 -> 10)", 38, [256, 256, 256], 10, 5, [4096, 512], [24576, 3072], "humanoid")
    rows += run("C4 Cassie + latent (187 -> 512-256-128 -> 12), T=8", 187, [512, 256, 128], 12, 8, [16384],
                [24576] if QUICK else [98304], "cassie", dobs=True)
    rows += run("RMA actual (338 -> 512-256-128 -> 16)", 338, [512, 256, 128], 16, 5, [4096], [24576], None, dobs=True)
    Bs = [1024, 16384] if QUICK else [1024, 4096, 16384, 65536, 262144, 1048576]
    for T in ([5] if QUICK else [5, 10, 20]):
        rows += run(f"C5 sweep, C1 dims, T={T}", 48, [256, 256], 12, T, Bs, [b for b in Bs if b <= 262144])
    lines = ["| config | T | batch | pass | ms | samples/s | algorithmic TFLOP/s |", "|---|---:|---:|---|---:|---:|---:|"]
    for name, T, B, kind, ms, sps, tf in rows:
        if ms is None:
            lines.append(f"| {name} | {T} | {B} | {kind} | - | - | {tf} |")
        else:
            lines.append(f"| {name} | {T} | {B} | {kind} | {ms:.3f} | {sps:.3e} | {tf:.1f} |")
    text = "\n".join(lines)
    print(text)
    if OUT:
        with open(OUT, "w") as f:
            f.write(text + "\n")


if __name__ == "__main__":
    main()
