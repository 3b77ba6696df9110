#!/bin/bash
# usage: tools/gpurun_retry.sh <log> <timeout_s> [--gpus N] -- '<command>'   (retries while the pod answers "transient"/busy)
log=$1; shift; to=$1; shift
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $to "$@" > $log 2>&1
  rc=$?
  if grep -q "status=transient" $log || [ $rc -eq 3 ]; then sleep 60; continue; fi
  break
done
exit $rc
