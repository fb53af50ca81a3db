#!/bin/bash
# usage: gpurun_retry.sh <out-file> <gpurun args...>   -- retries while the pod answers "transient / busy" (nothing is charged)
out=$1; shift
for i in $(seq 1 40); do
    /usr/local/graft/bin/gpurun "$@" > "$out" 2>&1
    rc=$?
    if grep -q "status=transient" "$out" || [ $rc -eq 3 ]; then sleep 150; continue; fi
    break
done
echo "gpurun_retry: attempts=$i rc=$rc" >> "$out"
