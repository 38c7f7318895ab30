#!/bin/bash
# tools/search_variants.sh <variant names...>: k_round launch time and sims/s of the bench for library variants built by
# tools/build_search_variant.sh ("base" = the regular library)
for v in "$@"; do
    if [ "$v" = base ]; then unset SPRL_B200_LIB; else export SPRL_B200_LIB=sprl_b200/lib/variants/$v.so; fi
    python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1])
print('$v', 'sims/s', d['value'], 'ms_per_step', d['ms_per_step'], 'k_round_ms', d['roofline_search']['launch_ms'], 'evalnet_ms', d['roofline']['launch_ms'], 'clock', d['clocks']['sm_mhz'])"
done
