#!/bin/bash
# Retry wrapper around gpurun: exit code 3 (no box / slot free, nothing charged) is retried every 90 s.
#   tools/gpu.sh <timeout-seconds> [--gpus N] -- '<command>'
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
