set -x
timeout 900 python -m pytest tests/test_evalnet.py -m gpu -x -q 2>&1 | tail -3
python bench.py --no-e2e --no-cpu-baseline > gpurun_out/bench_i.json 2> gpurun_out/bench_i.err; echo bench rc=$?; tail -2 gpurun_out/bench_i.err; python -c "
import json
d=json.load(open('gpurun_out/bench_i.json'))
print(d['value'], d['moves_per_sec'], d['roofline']['launch_ms'], d['roofline']['traffic'], d['roofline_search']['launch_ms'], d['roofline_search']['traffic'], d['roofline_search']['algorithmic_bytes_per_launch'])
"
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 120 --csv --log-file gpurun_out/launches_r1i.csv python bench.py --steps 1 --warmup 3 --rounds 16 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches_r1i.log 2>&1
