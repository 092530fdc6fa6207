"""Run the actor forward several times on the same input and localise run-to-run differences per layer
(debugging aid for pipeline races).  Usage: python tools/diag_repro.py [B] [runs]"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from fullysnnforleggedrobots_b200 import PopSpikeActor
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 24576
    runs = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    torch.manual_seed(0)
    actor = PopSpikeActor(48, 12, 10, 10, [256, 256], (-5, 5), math.sqrt(0.15), spike_ts=5).cuda()
    layers = actor.snn.linear_layers()
    prm = (actor.encoder.mean.data, actor.encoder.std.data, [l.weight.data for l in layers], [l.bias.data for l in layers],
           actor.decoder.decoder.weight.data, actor.decoder.decoder.bias.data)
    g = torch.Generator(device="cuda").manual_seed(1)
    obs = torch.randn(B, 48, device="cuda", generator=g)
    if os.environ.get("DIAG_BWD"):
        # forward + backward: compare every parameter gradient bitwise across runs
        G = torch.randn(B, 12, device="cuda", generator=g)
        gref = None
        for r in range(runs):
            for p_ in actor.parameters():
                p_.grad = None
            mu, _ = actor(obs)
            (mu * G).sum().backward()
            torch.cuda.synchronize()
            gr = {n: p_.grad.detach().clone() for n, p_ in actor.named_parameters() if p_.grad is not None}
            if gref is None:
                gref = gr
                continue
            for n in gr:
                d = (gr[n] != gref[n])
                if d.any():
                    rel = ((gr[n] - gref[n]).norm() / gref[n].norm()).item()
                    print(f"run {r}: grad {n}: {d.sum().item()} of {d.numel()} differ, rel {rel:.3e}")
            print(f"run {r}: done")
        return
    with torch.inference_mode():
        for _ in range(3):
            actor(obs)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            actor(obs)
        e1.record()
        torch.cuda.synchronize()
        print(f"forward (inference, incl. weight pack): {e0.elapsed_time(e1) / 10:.3f} ms")
    ref = None
    for r in range(runs):
        mu, trace = actor.engine.forward(obs, *prm, train=True)
        torch.cuda.synchronize()
        tr = actor.engine.unpack_trace(trace, B)
        if ref is None:
            ref = (mu.cpu(), tr)
            continue
        print(f"run {r}: mu differs in {(mu.cpu() != ref[0]).sum().item()} entries")
        names = ["enc"] + [f"L{i}" for i in range(len(tr["spikes"]))]
        for name, a, b in zip(names, [tr["enc_spikes"]] + tr["spikes"], [ref[1]["enc_spikes"]] + ref[1]["spikes"]):
            d = (a != b)
            print(f"  spikes {name}: {d.sum().item()} differ of {d.numel()}")
            if d.any():
                idx = d.nonzero()[:12]
                print("    first (t, b, n):", [tuple(int(x) for x in i) for i in idx], " tiles:", sorted({int(i[1]) // 48 for i in d.nonzero()})[:20])
        for i, (a, b) in enumerate(zip(tr["volt"], ref[1]["volt"])):
            d = (a != b)
            print(f"  volt L{i}: {d.sum().item()} differ")
            if d.any():
                nz = d.nonzero()
                print("    first (t, b, n):", [tuple(int(x) for x in i) for i in nz[:12]])
                print("    tiles:", sorted({int(i[1]) // 48 for i in nz})[:30], " neurons/128:", sorted({int(i[2]) // 128 for i in nz}),
                      " t:", sorted({int(i[0]) for i in nz}))
                t0 = int(nz[0][1]) // 48
                sel = nz[(nz[:, 1] // 48) == t0]
                print(f"    tile {t0}: samples-in-tile", sorted({int(x) % 48 for x in sel[:, 1]}), " n-range", int(sel[:, 2].min()), int(sel[:, 2].max()),
                      " count", len(sel))
                for tt in sorted({int(x) for x in sel[:, 0]}):
                    m = sel[sel[:, 0] == tt]
                    dv = (a[m[:, 0], m[:, 1], m[:, 2]] - b[m[:, 0], m[:, 1], m[:, 2]]).abs()
                    print(f"      t={tt}: {len(m)} entries, max|dv|={dv.max().item():.3e} median={dv.median().item():.3e}")


if __name__ == "__main__":
    main()
