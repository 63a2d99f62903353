#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout> '<command>'   -- retries while the pod answers "transient" (nothing charged)
T=$1; shift
for i in $(seq 1 12); do
  OUT=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1)
  echo "$OUT" | tail -25
  if ! echo "$OUT" | grep -q "status=transient"; then exit 0; fi
  echo "[retry $i] pod busy, sleeping 150 s"
  sleep 150
done
exit 3
