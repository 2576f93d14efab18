#!/bin/bash
# gpurun with retries while the pod has no free slot (exit code 3 = nothing charged).
# usage: tools/gpurun_retry.sh <logfile> <gpurun args...>
LOG=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "gpurun exit $rc (attempt $i)" >> "$LOG"; exit $rc; fi
  sleep 90
done
echo "gpurun: no slot after 40 attempts" >> "$LOG"; exit 3
