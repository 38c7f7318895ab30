#!/bin/bash
# tools/ab_variants.sh <variant names...>: the evaluator under sustained load (400 back-to-back forwards with NVML sampling)
# and the bench's device-timed line for library variants built by tools/build_variant.sh ("base" = the regular library).
for v in "$@"; do
    if [ "$v" = base ]; then unset SPRL_B200_LIB; else export SPRL_B200_LIB=sprl_b200/lib/variants/$v.so; fi
    echo "== $v"; REPS=400 python tools/check_resident.py 61960 2>&1 | grep -E "rows|resident: sm clock"
done
for v in "$@"; do
    if [ "$v" = base ]; then unset SPRL_B200_LIB; else export SPRL_B200_LIB=sprl_b200/lib/variants/$v.so; fi
    python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1])
print('$v', 'sims/s', d['value'], 'ms_per_step', d['ms_per_step'], 'evalnet_ms', d['roofline']['launch_ms'], 'clock', d['clocks']['sm_mhz'], 'W', d['clocks']['power_w'])"
done
