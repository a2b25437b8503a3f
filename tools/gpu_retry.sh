#!/bin/bash
# usage: [GPUS=2|4|8] tools/gpu_retry.sh <timeout_s> '<command>'  -- retries gpurun while the pod answers "busy" (nothing is charged)
T=$1; shift
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun ${GPUS:+--gpus $GPUS} --timeout "$T" -- "$@" > /tmp/gpu_retry_last.log 2>&1
  rc=$?
  if grep -q "status=transient" /tmp/gpu_retry_last.log; then sleep 120; continue; fi
  cat /tmp/gpu_retry_last.log; exit $rc
done
cat /tmp/gpu_retry_last.log; exit 3
