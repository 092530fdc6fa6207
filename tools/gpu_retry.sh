#!/bin/bash
# usage: tools/gpu_retry.sh TIMEOUT_S 'command'   — retries gpurun while the pod answers busy (rc 3 / transient)
T=$1; shift
for i in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient"; then sleep 90; continue; fi
  if [ $rc -eq 3 ]; then sleep 90; continue; fi
  echo "$out" | tail -60; exit $rc
done
echo "gave up after 30 tries"; exit 3
