#!/bin/bash
# retry gpurun while it answers "busy" (exit code 3: nothing charged); usage: tools/gpurun_retry.sh [gpurun options] -- 'command'
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3
