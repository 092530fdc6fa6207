#!/bin/bash
TAG=${1:-head}
mkdir -p gpurun_out
python tools/profile_head.py 24576 20 2>&1 | tail -1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"actor_head" -c 1 -f -o gpurun_out/${TAG}_head python tools/profile_head.py 24576 1 > gpurun_out/${TAG}_ncu_head.log 2>&1; echo "ncu rc $?"
