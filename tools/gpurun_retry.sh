#!/bin/bash
# Local helper: call gpurun until it gets a box (exit code 3 = nothing free right now, nothing charged).
# usage: tools/gpurun_retry.sh <log> <gpurun args...>
log=$1; shift
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 120
done
exit 3
