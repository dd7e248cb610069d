#!/bin/bash
# usage: gpu_retry.sh <logfile> <timeout> <command...>   -- retries while the pod answers busy (rc 3)
log=$1; shift; to=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  rc=$?
  if grep -q "status=transient" $log; then sleep 90; continue; fi
  exit $rc
done
