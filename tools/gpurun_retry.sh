#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout> <logfile> <command...>   — retries while the pod answers "busy" (exit code 3)
T=$1; LOG=$2; shift 2
for attempt in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@" > $LOG 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" $LOG; then exit $rc; fi
  sleep 90
done
exit 3
