#!/bin/bash
# usage: tools/gpu/retry.sh <log> <gpurun args...>; retries while gpurun answers "busy" (exit 3) or refuses transiently (2)
log=$1; shift
for i in $(seq 1 20); do
  gpurun "$@" > "$log" 2>&1; rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$log"; then exit $rc; fi
  sleep 240
done
exit 3
