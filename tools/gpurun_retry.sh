#!/bin/bash
# Retries a gpurun call while the pod answers "no slot" (exit code 3).  Usage: tools/gpurun_retry.sh <timeout_s> <command string>
T="$1"; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 100
done
exit 3
