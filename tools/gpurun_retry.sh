#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> '<command>' [extra gpurun args]
# Retries while gpurun answers "busy" (exit 3: nothing charged); any other exit code is final.
T=$1; shift
CMD=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" "$@" -- "$CMD"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
