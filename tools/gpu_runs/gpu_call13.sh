set -x
timeout 900 python -m pytest tests/test_evalnet.py -m gpu -x -q 2>&1 | tail -3
SPRL_EVALNET_TIMING=1 timeout 120 python tools/check_evalnet.py 32768 2 2>&1 | grep "forward B\|CTA 0\|dlogit" | tail -3
for sl in 8192 16384 32768; do
python bench.py --slots $sl --no-e2e --no-cpu-baseline > gpurun_out/bench_g_$sl.json 2> gpurun_out/bench_g_$sl.err; echo bench rc=$?; tail -2 gpurun_out/bench_g_$sl.err; python -c "
import json,sys
d=json.load(open('gpurun_out/bench_g_$sl.json'))
print($sl, d['value'], d['moves_per_sec'], d['roofline']['launch_ms'], d['roofline']['leaves_per_launch'], d['roofline']['issued_frac'], d['roofline_search']['launch_ms'])
"
done
