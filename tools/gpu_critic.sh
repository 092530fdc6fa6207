#!/bin/bash
TAG=${1:-critic}
mkdir -p gpurun_out
python tools/profile_critic.py 24576 20 2>&1 | tail -1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"scan_gemm" -c 3 -f -o gpurun_out/${TAG}_critic python tools/profile_critic.py 24576 1 > gpurun_out/${TAG}_ncu_critic.log 2>&1; echo "ncu rc $?"
