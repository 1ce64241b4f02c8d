#!/bin/bash
# usage: tools/gpu_retry.sh <timeout-seconds> [--gpus N] '<command>'   -- retries while the pod has no free GPU slot (exit code 3)
T=$1; shift
G=""
if [ "$1" = "--gpus" ]; then G="--gpus $2"; shift 2; fi
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" $G -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
