#!/bin/bash
# usage: scratch/gpu_retry.sh <timeout> <logfile> <command...>   — retries while the pod answers busy (exit 3)
T=$1; LOG=$2; shift 2
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@" > $LOG 2>&1
  rc=$?
  if grep -q "status=transient" $LOG || [ $rc -eq 3 ]; then sleep 60; continue; fi
  exit $rc
done
exit 3
