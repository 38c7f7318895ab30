# k_heads: K-split over 512 threads (16 warps per SM)
set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for b in 32768 65536; do timeout 100 python tools/check_evalnet.py $b 2 2>&1 | grep "forward B\|rror" | tail -2; done
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_heads3.csv python tools/check_evalnet.py 65536 2 > /dev/null 2>&1; python tools/ncu_summary.py launches gpurun_out/launches_heads3.csv | head -5
python bench.py --no-e2e --no-cpu-baseline --steps 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('sims/s %.2f M  k_round %.4f ms  evalnet %.4f ms' % (d['value']/1e6, d['roofline_search']['launch_ms'], d['roofline']['launch_ms']))"
