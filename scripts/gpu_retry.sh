#!/bin/bash
# Usage: scripts/gpu_retry.sh LOGFILE [gpurun args...] -- 'command'
# Retries a gpurun call while the pod answers "busy / draining" (exit code 3: nothing charged), every 2 minutes.
log=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
