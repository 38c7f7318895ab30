set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 900 python tools/bench_configs.py > gpurun_out/bench_configs_r1j.jsonl 2> gpurun_out/bench_configs_r1j.err; echo rc=$?; tail -3 gpurun_out/bench_configs_r1j.err; grep "selfplay\|match" gpurun_out/bench_configs_r1j.jsonl
