#!/bin/bash
# dX kernels: depth of the g_c unit ring (tools build, -D switch)
for v in 4 5 4 5; do
  echo "SNN_DX_NS=$v"
  SNN_EXTRA_NVCC="-DSNN_DX_NS=$v" SNN_DBG=512 python tools/bench_kernels.py 24576 2>&1 | grep -E " dx "
done
