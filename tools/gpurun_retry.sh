#!/usr/bin/env bash
# gpurun with retries on "busy" (exit 3: nothing charged).  usage: tools/gpurun_retry.sh <log> <gpurun args...>
log=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
