#!/bin/bash
# usage: tools/gpurun_retry.sh <logfile> <gpurun args...>   -- retries while the pod answers "busy" (nothing is charged for those)
LOG=$1; shift
for i in $(seq 1 40); do
  gpurun "$@" > "$LOG" 2>&1
  if grep -q "status=transient\|backing off" "$LOG"; then sleep 90; continue; fi
  break
done
