# k_round instruction diet (next_warp, shuffle prior sum): parity + bench
set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --no-e2e --no-cpu-baseline --steps 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('sims/s %.2f M  k_round %.4f ms  evalnet %.4f ms' % (d['value']/1e6, d['roofline_search']['launch_ms'], d['roofline']['launch_ms']))"
