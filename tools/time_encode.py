"""GPU-only time of the inference forward (SNN_DBG=64: encoder kernel alone), via CUDA-graph replay."""
import math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
# ablation switches / cycle counters live in the -DSNN_ABLATE tools build only (never in the release library)
if os.environ.get("SNN_DBG", "0") != "0":
    os.environ["SNN_LIB"] = "ablate"
    from fullysnnforleggedrobots_b200 import _lib as _l
    _l.build(ablate=True)
from fullysnnforleggedrobots_b200 import PopSpikeActor
B = int(sys.argv[1]) if len(sys.argv) > 1 else 24576
actor = PopSpikeActor(48, 12, 10, 10, [256, 256], (-5, 5), math.sqrt(0.15), spike_ts=5).cuda()
obs = torch.randn(B, 48, device="cuda")
with torch.inference_mode():
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            actor(obs)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10):
            actor(obs)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
print(f"B={B} SNN_DBG={os.environ.get('SNN_DBG','0')}: {e0.elapsed_time(e1)/100*1e3:.1f} us per call (graph replay)")
