#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout> '<command>'   -- retries while the pod is busy
for i in $(seq 1 20); do
  out=$(gpurun --timeout "$1" -- "$2" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 90; continue; fi
  echo "$out" | grep -v "^\[gpurun\] sending"
  exit 0
done
echo "gave up: pod busy"; exit 3
