#!/bin/bash
# usage: tools/gpurun_retry.sh <log> <gpurun args...>   -- repeats the call while the pod answers "busy" (exit 3 / transient)
log=$1; shift
for i in $(seq 1 15); do
    /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
    rc=$?
    if grep -q "status=transient" "$log" || [ $rc -eq 3 ]; then sleep 150; continue; fi
    exit $rc
done
exit 3
