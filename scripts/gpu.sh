#!/bin/bash
# gpurun with retries while the pod answers "busy / draining" (exit code 3, nothing charged).
#   scripts/gpu.sh <timeout-seconds> [--gpus N] -- '<command>'
T=$1; shift
for i in $(seq 1 40); do
    /usr/local/graft/bin/gpurun --timeout "$T" "$@" > /tmp/gpu_sh_last.txt 2>&1
    rc=$?
    if [ $rc -ne 3 ] && ! grep -q "status=transient" /tmp/gpu_sh_last.txt; then
        cat /tmp/gpu_sh_last.txt
        exit $rc
    fi
    sleep 45
done
cat /tmp/gpu_sh_last.txt
exit 3
