#!/bin/bash
# usage: tools/gpurun_retry.sh <log> <gpurun args...> : retries while the pod answers "no slot right now" (exit code 3, nothing charged)
log=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1; rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 90
done
exit 3
