#!/bin/bash
mkdir -p build
for v in "3 4" "7 1" "12 1" "3 1" "6 2" "4 3"; do set -- $v
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DKSLOTS=$1 -DKWARPS=$2 -I fullysnnforleggedrobots_b200/csrc tools/l2_bulk_rate.cu -o build/l2_bulk_rate_$1_$2 || exit 1
  echo "slots per warp $1, issuing warps $2"; ./build/l2_bulk_rate_$1_$2 | grep -E "^16384|^8192"
done
