#!/bin/bash
# tools/gpu.sh [--gpus N] <timeout-seconds> '<command>' -- sends the repo to a B200 box with gpurun, retrying while no
# box or slot is free (exit code 3 = nothing charged).
GP=""
if [ "$1" = "--gpus" ]; then GP="--gpus $2"; shift 2; fi
T=$1; shift
for i in $(seq 1 40); do
    /usr/local/graft/bin/gpurun $GP --timeout "$T" -- "$@"
    rc=$?
    if [ $rc -ne 3 ]; then exit $rc; fi
    sleep 90
done
exit 3
