#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3): tools/gpu_retry.sh TIMEOUT 'command' [gpurun flags]
T=$1; CMD=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" --timeout "$T" -- "$CMD"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 60
done
exit 3
