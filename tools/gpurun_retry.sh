#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-s> <log> <command...>   -- retries while the pod answers "no box" (exit 3)
T=$1; LOG=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@" > $LOG 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "gpurun rc=$rc after $i tries" >> $LOG; exit $rc; fi
  sleep 120
done
