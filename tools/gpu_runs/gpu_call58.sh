# branch-free sym_cell (k_round, k_emit), distributions written by the state loop of k_emit: parity, ncu of k_emit, bench
set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_emit' -c 2 \
  -o gpurun_out/prof_emit_r1u -f python tools/profile_env.py 6 12 2048 > gpurun_out/ncu_emit_r1u.log 2>&1; echo ncu emit rc=$?
python bench.py --no-e2e --no-cpu-baseline --steps 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('sims/s %.2f M  k_round %.4f ms  evalnet %.4f ms' % (d['value']/1e6, d['roofline_search']['launch_ms'], d['roofline']['launch_ms']))"
