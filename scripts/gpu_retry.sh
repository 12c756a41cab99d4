#!/bin/bash
# usage: scripts/gpu_retry.sh LOG TIMEOUT [--gpus N] -- 'command'   (retries while the pod has no free slot: exit code 3)
log=$1; shift; to=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$to" "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
