#!/bin/bash
# same-box A/B/A/B of one environment switch of the fused PPO step: tools/gpu_env_ab.sh TAG VAR
TAG=${1:-envab}; VAR=${2:-SNN_UPDATE_GRAPH}
mkdir -p gpurun_out
for v in 1 0 1 0 1 0; do
  env $VAR=$v timeout 300 python bench.py --quick --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$VAR=$v ms_per_step', round(d['ms_per_step'],3))"
done | tee gpurun_out/${TAG}_${VAR}_ab.txt
