#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout> '<command>'
# Retries ONLY when gpurun answers exit code 3 (no box/slot free: nothing charged, nothing ran).
# Every other outcome -- including "transient"/lost-box verdicts, which count as strikes -- is
# printed in full and returned: never retry those blindly, read the verdict first.
for i in $(seq 1 20); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$1" -- "$2" 2>&1)
  rc=$?
  if [ "$rc" = "3" ]; then sleep 90; continue; fi
  echo "$out" | grep -v "^\[gpurun\] sending"
  exit $rc
done
echo "gave up: pod busy"; exit 3
