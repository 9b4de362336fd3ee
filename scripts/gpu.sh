#!/bin/bash
# gpurun with retries while the pod answers "busy / draining" (exit code 3, nothing charged).
#   scripts/gpu.sh <timeout-seconds> [--gpus N] -- '<command>'
# Before every (re)submission the tree must be self-consistent (the snapshot is taken at submission time): the library has
# to load with every symbol the Python ABI table names.
T=$1; shift
for i in $(seq 1 60); do
    until python -c "from qwen3_rs_b200 import transformer as T; T.load_library()" >/dev/null 2>&1; do sleep 5; done
    /usr/local/graft/bin/gpurun --timeout "$T" "$@" > /tmp/gpu_sh_last_$$.txt 2>&1
    rc=$?
    if [ $rc -ne 3 ] && ! grep -q "status=transient" /tmp/gpu_sh_last_$$.txt; then
        cat /tmp/gpu_sh_last_$$.txt
        exit $rc
    fi
    sleep 45
done
cat /tmp/gpu_sh_last_$$.txt
exit 3
