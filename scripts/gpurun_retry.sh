#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout-seconds> '<command>'  -- re-issues a gpurun call while the pod answers "busy" (nothing charged)
T=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@" > /tmp/gpurun_retry.out 2>&1
  if grep -q '"status": "transient"' gpurun_out/.last_call.json 2>/dev/null; then sleep 90; continue; fi
  break
done
tail -40 /tmp/gpurun_retry.out
