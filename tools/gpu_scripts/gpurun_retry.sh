#!/bin/bash
# gpurun with retries while the pod answers "busy / draining" (exit 3, nothing charged).  usage: gpurun_retry.sh <log> <gpurun args...>
LOG=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  rc=$?
  if grep -q "status=transient" "$LOG" || [ $rc -eq 3 ]; then sleep 45; continue; fi
  exit $rc
done
exit 3
