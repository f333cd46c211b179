#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> '<command>' [extra gpurun flags]   — retries while the pod answers "transient" (exit 3)
T=$1; CMD=$2; shift 2
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@" --timeout "$T" -- "$CMD"; rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
