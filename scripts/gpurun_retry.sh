#!/bin/bash
# usage: scripts/gpurun_retry.sh <log> <timeout_s> [--gpus N] -- <command...>
# retries a gpurun call while the pod answers "busy" (exit code 3, nothing charged), then stops
LOG=$1; shift
TMO=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$TMO" "$@" > "$LOG" 2>&1
  rc=$?
  echo "attempt $i rc=$rc" >> "$LOG.attempts"
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
