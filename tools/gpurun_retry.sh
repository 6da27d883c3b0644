#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> '<command>'   [GPUS=n]
# retries while the pod answers busy (status=transient, nothing charged)
T=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun ${GPUS:+--gpus $GPUS} --timeout $T -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient rc=None charged=0.0s"; then
    sleep 90
    continue
  fi
  echo "$out"
  exit 0
done
echo "gpurun_retry: still busy after 40 attempts"
exit 3
