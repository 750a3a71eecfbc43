#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout> <logfile> <command...>   — retries while the pod answers "transient" (nothing charged)
T=$1; LOG=$2; shift 2
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@" > $LOG 2>&1
  if ! grep -q "status=transient" $LOG; then break; fi
  sleep 90
done
tail -n 5 $LOG
