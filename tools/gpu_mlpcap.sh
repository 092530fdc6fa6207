#!/bin/bash
# A/B of the critic's CTA cap (SNN_MLP_MAX_CTAS) with dynamic tile scheduling in the actor's kernels
for c in 0 111 74 48 0; do
  SNN_MLP_MAX_CTAS=$c python bench.py --quick --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('SNN_MLP_MAX_CTAS=$c ms_per_step', round(d['ms_per_step'],3))"
done
