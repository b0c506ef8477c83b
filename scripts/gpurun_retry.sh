#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout-seconds> '<command>' [extra gpurun flags...]
# Retries a gpurun call while the pod answers "busy / draining" (nothing is charged for those).
T=$1; CMD=$2; shift 2
for i in $(seq 1 20); do
  OUT=$(/usr/local/graft/bin/gpurun "$@" --timeout "$T" -- "$CMD" 2>&1)
  if echo "$OUT" | grep -q "status=transient\|status=busy\|rc=3"; then
    echo "[retry $i] pod busy"; sleep 150; continue
  fi
  echo "$OUT"; exit 0
done
echo "gave up"; exit 3
