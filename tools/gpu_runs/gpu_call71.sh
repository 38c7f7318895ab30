# last build of the round: GPU tests, smoke, short bench
set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --no-cpu-baseline --steps 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('sims/s %.2f M  e2e %.2f M  k_round %.4f ms  evalnet %.4f ms' % (d['value']/1e6, d['e2e']['value']/1e6, d['roofline_search']['launch_ms'], d['roofline']['launch_ms']))"
