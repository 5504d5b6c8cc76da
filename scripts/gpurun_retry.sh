#!/bin/bash
# usage: gpurun_retry.sh <timeout> <logfile> <command...>: retries while the pod answers busy (rc 3 / transient)
T=$1; LOG=$2; shift 2
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@" > $LOG 2>&1
  if ! grep -q "status=transient\|nothing was charged" $LOG; then exit 0; fi
  sleep 150
done
