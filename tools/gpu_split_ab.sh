#!/bin/bash
# fused step: split head (actor_head_kernel on the actor's chain) vs the one-kernel head after a join, same box, A/B/A/B
TAG=${1:-split}
mkdir -p gpurun_out
for v in 1 0 1 0; do
  SNN_SPLIT_HEAD=$v timeout 300 python bench.py --quick --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('SNN_SPLIT_HEAD=$v ms_per_step', round(d['ms_per_step'],3))"
done | tee gpurun_out/${TAG}_split_ab.txt
