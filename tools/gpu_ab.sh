#!/bin/bash
# quick check of a kernel change: GPU suite, per-kernel times at C2, short bench, timeline
TAG=${1:-ab}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_gpu_tests.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/${TAG}_gpu_tests.log
timeout 200 python tools/bench_kernels.py 24576 > gpurun_out/${TAG}_kernels.txt 2>&1; cat gpurun_out/${TAG}_kernels.txt
timeout 200 python tools/bench_kernels.py 3072 48,256,256,256,12 >> gpurun_out/${TAG}_kernels.txt 2>&1
timeout 300 python bench.py --quick --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_quick.json 2> gpurun_out/${TAG}_bench_quick.err; echo "bench rc $?"
python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_bench_quick.json').read().strip().splitlines()[-1])
print('ms_per_step', d['ms_per_step'], 'value', d['value'])
print({k: round(v['us_per_launch'],1) for k,v in d['roofline']['per_layer'].items()})
"
timeout 200 python tools/step_timeline.py graph gpurun_out/${TAG}_update_timeline.md > gpurun_out/${TAG}_timeline.log 2>&1; echo "timeline rc $?"
sed -n 1,12p gpurun_out/${TAG}_update_timeline.md
