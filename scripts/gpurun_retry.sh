#!/bin/bash
# gpurun_retry.sh <log> <gpurun args...>: retries while the pod answers "transient" (no slot; nothing charged)
LOG=$1; shift
for attempt in $(seq 1 ${GPURUN_TRIES:-30}); do
  /usr/local/graft/bin/gpurun "$@" > $LOG 2>&1
  if grep -q "status=transient" $LOG; then sleep ${GPURUN_SLEEP:-90}; continue; fi
  break
done
echo "attempts: $attempt" >> $LOG
