#!/bin/bash
# forward epilogue A/B (tools build): voltage-code encoders, and without the trace stores (SNN_DBG=128)
TAG=${1:-encab}
mkdir -p gpurun_out
for v in "0 512" "1 512" "0 128" "0 512" "1 512"; do set -- $v
  echo "VOLT_ENC=$1 SNN_DBG=$2"
  SNN_EXTRA_NVCC="-DSNN_VOLT_ENC=$1" SNN_DBG=$2 python tools/bench_kernels.py 24576 2>&1 | grep -E " dx | fwd |total"
done > gpurun_out/${TAG}_encab.txt 2>&1
cat gpurun_out/${TAG}_encab.txt
