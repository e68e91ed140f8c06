#!/bin/bash
# Development helper: run one gpurun call, retrying while the pod answers "busy / draining" (exit code 3, nothing
# charged).  Usage: tools/gpurun_retry.sh <log file> <timeout s> '<command>'
log=$1; to=$2; shift 2
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$to" -- "$@" > "$log" 2>&1
  rc=$?
  if grep -q "status=transient" "$log" || [ $rc -eq 3 ]; then sleep 45; continue; fi
  break
done
echo "[gpurun_retry] attempts=$attempt rc=$rc" >> "$log"
