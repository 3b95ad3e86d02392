#!/bin/bash
# usage: [GPUS=8] tools/gpurun_retry.sh <log> <timeout_s> <command...>   -- retries while the pod answers busy (exit 3)
log=$1; shift; to=$1; shift
extra=""
[ -n "$GPUS" ] && extra="--gpus $GPUS"
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $to $extra -- "$@" > $log 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" $log; then exit $rc; fi
  sleep 120
done
exit 3
