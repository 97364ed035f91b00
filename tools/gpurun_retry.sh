#!/bin/bash
# gpurun with retries while the pod answers "transient" (no slot free; nothing charged).  usage: tools/gpurun_retry.sh [gpurun args] -- 'command'
for i in $(seq 1 40); do
    out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
    if echo "$out" | grep -q "status=transient"; then sleep 120; continue; fi
    echo "$out"
    exit 0
done
echo "gpurun_retry: gave up"; exit 3
