#!/bin/bash
# gpurun with retries while the pod answers "transient" (nothing charged): tools/gpurun_retry.sh <timeout> '<command>'
T=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@" > /tmp/gpurun_try.log 2>&1
  if grep -q "status=transient" /tmp/gpurun_try.log; then sleep 60; continue; fi
  break
done
cat /tmp/gpurun_try.log
