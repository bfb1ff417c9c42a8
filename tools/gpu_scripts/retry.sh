#!/bin/bash
# usage: retry.sh <timeout> '<command>'  -- repeats a gpurun call while the pod answers "busy" (nothing is charged for those)
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$1" -- "$2" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 60; continue; fi
  echo "$out"; exit 0
done
echo "$out"; exit 3
