#!/bin/bash
# usage: tools/gpurun_retry.sh <log> <timeout> <command...>   - retries while gpurun answers "no slot" (rc 3)
LOG=$1; shift; TMO=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $TMO -- "$@" > $LOG 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
