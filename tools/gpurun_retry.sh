#!/bin/bash
# usage: tools/gpurun_retry.sh <logfile> <gpurun args...>   — retries while the pod answers "transient" (nothing charged)
log=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  if grep -q "status=transient" "$log" || grep -q "retry in a few minutes" "$log"; then sleep 90; continue; fi
  break
done
tail -100 "$log"
