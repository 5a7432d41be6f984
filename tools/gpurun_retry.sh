#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> '<command>'
# Runs the command on a B200 box through gpurun and retries (nothing is charged) while the pod answers "busy".
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout "$1" -- "$2" > /tmp/gpurun_retry.out 2>&1
  if ! grep -q "status=transient" /tmp/gpurun_retry.out; then break; fi
  sleep 45
done
tail -60 /tmp/gpurun_retry.out
