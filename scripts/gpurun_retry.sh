#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout> <logfile> [--gpus N] <command>   -- retries while the pool answers "busy"
T=$1; LOG=$2; shift 2
EXTRA=""
if [ "$1" == "--gpus" ]; then EXTRA="--gpus $2"; shift 2; fi
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" $EXTRA -- "$@" > "$LOG" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$LOG"; then exit $rc; fi
  sleep 90
done
exit 3
