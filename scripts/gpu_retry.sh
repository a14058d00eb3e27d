#!/bin/bash
# usage: gpu_retry.sh <logfile> <timeout> <command...>  — retries while the pod answers "busy" (exit code 3 / transient)
LOG=$1; TO=$2; shift 2
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $TO -- "$@" > $LOG 2>&1
  if ! grep -q "status=transient" $LOG; then exit 0; fi
  sleep 60
done
