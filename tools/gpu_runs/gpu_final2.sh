set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py > gpurun_out/bench_h.json 2> gpurun_out/bench_h.err; echo bench rc=$?; tail -2 gpurun_out/bench_h.err; cat gpurun_out/bench_h.json
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_h_ref.json 2> gpurun_out/bench_h_ref.err; cat gpurun_out/bench_h_ref.json
timeout 900 python tools/bench_configs.py > gpurun_out/bench_configs_r1h.jsonl 2> gpurun_out/bench_configs_r1h.err; echo rc=$?; tail -3 gpurun_out/bench_configs_r1h.err; cat gpurun_out/bench_configs_r1h.jsonl
bash tools/worker_smoke.sh 2>&1 | tail -3
bash tools/evaluate_smoke.sh 2>&1 | tail -6
