#!/bin/bash
# usage: tools/gpurun_retry.sh <log> <gpurun args...>   -- retries while the pod answers "busy" / "draining" (nothing charged)
log=$1; shift
for i in $(seq 1 30); do
  gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$log"; then exit $rc; fi
  sleep 75
done
exit 3
