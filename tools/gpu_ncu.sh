#!/bin/bash
# ncu --set full of the tensor kernels of one training forward + backward (C2 minibatch) + a launch list of the bench
TAG=${1:-ncu}
mkdir -p gpurun_out
python tools/profile_step.py 24576 1 > /dev/null 2>&1 || { echo "profile_step failed without ncu"; exit 1; }
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"nm_gemm|dw_gemm" -c 8 -f -o gpurun_out/${TAG}_prof python tools/profile_step.py 24576 1 > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc $?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --quick --steps 1 --warmup 3 > gpurun_out/${TAG}_ncu_bench.log 2>&1; echo "ncu list rc $?"
ls -la gpurun_out | grep ${TAG}
