#!/bin/bash
# one GPU session: A/B of the asynchronous kernel's knobs (one process per configuration, each under its own timeout) + an ncu capture without the instrumented sections
export PROFILE_SITES=${PROFILE_SITES:-18944} PROFILE_DAYS=2
LOG=${LOG:-gpurun_out/r2_ab2.log}
for cfg in "$@"; do
  timeout 150 python tools/gpu_async_ab.py "$cfg" >> $LOG 2>&1 || echo "== $cfg : FAILED or timed out (rc $?)" >> $LOG
done
cat $LOG
