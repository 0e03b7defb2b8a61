#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> '<command>'   -- retries while the pod answers busy/transient (exit 3 or "transient")
T=$1; shift
for i in $(seq 1 12); do
  out=$(/usr/local/graft/bin/gpurun ${GPURUN_GPUS:+--gpus $GPURUN_GPUS} --timeout "$T" -- "$@" 2>&1)
  rc=$?
  echo "$out" | tail -60
  if echo "$out" | grep -q "status=transient\|status=busy" || [ $rc -eq 3 ]; then sleep 120; continue; fi
  exit $rc
done
exit 3
