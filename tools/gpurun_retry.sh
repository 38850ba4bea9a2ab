#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit 3: nothing charged).  usage: tools/gpurun_retry.sh LOG [gpurun args...]
LOG=$1; shift
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 150
done
exit 3
