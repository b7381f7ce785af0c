#!/bin/bash
# tools/gpu.sh <tag> <timeout_s> <command...>: one gpurun call with retries while the pod answers "busy";
# the call log goes to gpurun_out/<tag>_call.log
tag=$1; to=$2; shift 2
mkdir -p gpurun_out
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun ${GPUS:+--gpus $GPUS} --timeout "$to" -- "$@" > gpurun_out/${tag}_call.log 2>&1
  rc=$?
  if grep -q "status=transient\|status=busy\|rc=None" gpurun_out/${tag}_call.log && [ $rc -ne 0 ]; then sleep 45; continue; fi
  break
done
tail -40 gpurun_out/${tag}_call.log
