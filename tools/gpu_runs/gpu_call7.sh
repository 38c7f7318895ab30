set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python tools/bench_configs.py > gpurun_out/bench_configs_r1e.jsonl 2> gpurun_out/bench_configs_r1e.err; echo rc=$?; tail -3 gpurun_out/bench_configs_r1e.err; cat gpurun_out/bench_configs_r1e.jsonl
