#!/bin/bash
# dX epilogue A/B (tools build): voltage-code load depth x L2 prefetch distance
TAG=${1:-dxab}
mkdir -p gpurun_out
for v in "2 0" "2 1" "3 0" "3 1"; do set -- $v
  echo "VDEPTH=$1 PF_AHEAD=$2"
  SNN_EXTRA_NVCC="-DSNN_DX_VDEPTH=$1 -DSNN_DX_PF_AHEAD=$2" SNN_DBG=512 python tools/bench_kernels.py 24576 2>&1 | grep -E " dx | fwd |total"
done > gpurun_out/${TAG}_dxab.txt 2>&1
cat gpurun_out/${TAG}_dxab.txt
