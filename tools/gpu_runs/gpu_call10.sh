set -x
timeout 900 python -m pytest tests/test_evalnet.py -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_e.json 2> gpurun_out/bench_e.err; echo bench rc=$?; tail -2 gpurun_out/bench_e.err; cat gpurun_out/bench_e.json
