#!/bin/bash
# usage: scripts/gpurun_retry.sh <outfile> <gpurun args...>   -- retries while the pod has no free slot (exit code 3)
out=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@" > "$out" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "exit $rc" >> "$out"; exit $rc; fi
  sleep 45
done
echo "gave up" >> "$out"
