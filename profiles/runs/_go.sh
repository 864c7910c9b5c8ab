#!/bin/bash
# usage: profiles/_go.sh <log> <timeout> <command...>   — gpurun with retries while the pod is busy (exit code 3)
log=$1; shift; to=$1; shift
for i in $(seq 1 40); do
  gpurun --timeout $to -- "$@" > $log 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "gpurun rc=$rc (attempt $i)" >> $log; exit $rc; fi
  sleep 60
done
