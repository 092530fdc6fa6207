#!/bin/bash
# forward kernel: weight-ring depth vs operand slots (tools build, -D switches)
for v in "7 3" "9 2" "8 2" "6 4" "5 4"; do set -- $v
  echo "NW=$1 NS=$2"
  SNN_EXTRA_NVCC="-DSNN_FWD_NW=$1 -DSNN_FWD_NS=$2" SNN_DBG=512 python tools/bench_kernels.py 24576 2>&1 | grep -E " fwd "
done
