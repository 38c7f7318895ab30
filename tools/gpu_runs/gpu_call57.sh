# k_emit with a (games, moves) grid: sample parity + ncu; bench at 32,768 concurrent games
set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_emit' -c 2 \
  -o gpurun_out/prof_emit_r1t -f python tools/profile_env.py 6 12 2048 > gpurun_out/ncu_emit_r1t.log 2>&1; echo ncu emit rc=$?
timeout 600 python bench.py --slots 32768 --no-e2e --no-cpu-baseline --steps 2 > gpurun_out/bench_32k.json 2> gpurun_out/bench_32k.err; echo rc=$?
python -c "
import json
d=json.load(open('gpurun_out/bench_32k.json'))
print('32k slots: sims/s %.2f M  k_round %.4f ms  evalnet %.4f ms' % (d['value']/1e6, d['roofline_search']['launch_ms'], d['roofline']['launch_ms']))"
timeout 300 python tools/explore_overlap.py 16384 2>&1 | grep "groups="
