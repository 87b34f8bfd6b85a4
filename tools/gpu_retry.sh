#!/bin/bash
# usage: tools/gpu_retry.sh TIMEOUT [--gpus N] -- 'command'   : retries gpurun while it answers "busy" (exit 3)
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
