#!/bin/bash
# usage: tools/gpu_retry.sh <timeout> '<command>'  - retries while the pod answers "busy" (exit code 3)
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout "$1" -- "$2"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 75
done
exit 3
