#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout> <command...>   -- retries while the pod has no free slot (exit code 3)
T=$1; shift
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun --timeout $T "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 100
done
exit 3
