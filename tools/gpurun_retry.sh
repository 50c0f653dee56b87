#!/bin/bash
# Usage: tools/gpurun_retry.sh <timeout-seconds> <gpus> '<command>'  -- retries while the pod has no box free (rc 3 / transient)
t=$1; g=$2; shift 2
for attempt in 1 2 3 4 5 6 7 8; do
  if [ "$g" = "1" ]; then /usr/local/graft/bin/gpurun --timeout "$t" -- "$@" > /tmp/gpurun_last.out 2>&1; else /usr/local/graft/bin/gpurun --gpus "$g" --timeout "$t" -- "$@" > /tmp/gpurun_last.out 2>&1; fi
  rc=$?
  if grep -q "status=transient" /tmp/gpurun_last.out || [ $rc -eq 3 ]; then echo "attempt $attempt: no box, retrying in 90 s"; sleep 90; continue; fi
  break
done
tail -60 /tmp/gpurun_last.out
