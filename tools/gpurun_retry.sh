#!/bin/bash
# usage: [GPUS=N] tools/gpurun_retry.sh <timeout-seconds> <command string>   (retries while the pod answers "busy", rc 3)
t=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun ${GPUS:+--gpus $GPUS} --timeout "$t" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
