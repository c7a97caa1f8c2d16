#!/bin/bash
# usage: tools/gpurun_retry.sh <logfile> <gpurun args...>   -- retries while the pod answers "busy" (exit 3, nothing charged)
LOG=$1; shift
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$LOG"; then echo "gpurun rc=$rc (attempt $attempt)" >> "$LOG"; exit $rc; fi
  sleep 90
done
echo "gave up after 40 attempts" >> "$LOG"; exit 3
