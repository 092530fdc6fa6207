"""Run the layered inference forward several times on the same input; print which rows / batch tiles of mu differ
between runs (debugging aid for the tile scheduler).  Usage: python tools/diag_infer.py [B] [runs]"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fullysnnforleggedrobots_b200 import PopSpikeActor

B = int(sys.argv[1]) if len(sys.argv) > 1 else 24576
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 6
torch.manual_seed(0)
actor = PopSpikeActor(48, 12, 10, 10, [256, 256], (-5, 5), math.sqrt(0.15), spike_ts=5).cuda()
g = torch.Generator(device="cuda").manual_seed(1)
obs = torch.randn(B, 48, device="cuda", generator=g)
layers = actor.snn.linear_layers()
prm = (actor.encoder.mean.data, actor.encoder.std.data, [l.weight.data for l in layers], [l.bias.data for l in layers],
       actor.decoder.decoder.weight.data, actor.decoder.decoder.bias.data)
mu_t, _ = actor.engine.forward(obs, *prm, train=True)
ref = mu_t.clone()
with torch.inference_mode():
    for r in range(runs):
        mu, _ = actor(obs)
        torch.cuda.synchronize()
        d = (mu != ref).any(dim=1)
        rows = d.nonzero().flatten().cpu()
        tiles = sorted({int(x) // 48 for x in rows})
        print(f"run {r}: {int(d.sum())} rows differ from the training-mode forward; tiles {tiles[:24]}{' ...' if len(tiles) > 24 else ''} ({len(tiles)} tiles)")
        if len(rows):
            t0 = tiles[0]
            print("   rows of the first tile:", [int(x) % 48 for x in rows if int(x) // 48 == t0])
