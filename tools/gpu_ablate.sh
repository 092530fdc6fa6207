#!/bin/bash
# where does the forward kernel's time go: barrier-wait cycles per role + work-skipping ablations (tools build only)
TAG=${1:-abl}; shift
MASKS=${*:-0 1 2 3 4}
mkdir -p gpurun_out
python tools/role_cycles.py 24576 > gpurun_out/${TAG}_role_cycles.txt 2>&1; cat gpurun_out/${TAG}_role_cycles.txt | head -70
for m in $MASKS; do SNN_DBG=$m python tools/bench_kernels.py 24576 2>&1 | grep -E "SNN_DBG|fwd|dx " ; done > gpurun_out/${TAG}_ablations.txt 2>&1
cat gpurun_out/${TAG}_ablations.txt
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
