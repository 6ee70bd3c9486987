#!/bin/bash
# usage: tools/gpu_retry.sh <timeout-seconds> '<command>'   -- retries gpurun while the pod answers "transient"/busy (rc 3)
T=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient\|status=busy\|no box"; then sleep 90; continue; fi
  echo "$out"; exit 0
done
echo "gave up after 40 tries"; echo "$out" | tail -5
