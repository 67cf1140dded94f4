#!/bin/bash
# usage: tools/gpurun_retry.sh [gpurun args...] -- 'command'   (retries while the pod answers "busy", exit code 3)
for i in 1 2 3 4 5 6 7 8; do
  gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3
