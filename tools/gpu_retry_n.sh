#!/bin/bash
# usage: tools/gpu_retry_n.sh <gpus> <timeout_s> '<command>' : gpurun --gpus N with busy-retry
n=$1; t=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --gpus "$n" --timeout "$t" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
