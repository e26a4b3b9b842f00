#!/bin/bash
# gpurun with retries while the pod answers "transient" (nothing charged): usage gpurun_retry.sh <tries> <gpurun args...>
tries=$1; shift
for i in $(seq 1 $tries); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
  echo "$out" | tail -60
  if echo "$out" | grep -q "status=transient\|rc=3\b"; then echo "[retry $i/$tries in 100 s]"; sleep 100; continue; fi
  break
done
