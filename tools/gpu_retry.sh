#!/bin/bash
# usage: tools/gpu_retry.sh <logfile> <timeout> <command...>   -- retries gpurun while the pod answers busy (nothing charged)
log=$1; shift; to=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  if grep -q "status=transient\|status=busy\|rc=3" $log; then sleep 60; continue; fi
  break
done
