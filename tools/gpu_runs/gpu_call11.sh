set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py > gpurun_out/bench_f.json 2> gpurun_out/bench_f.err; echo bench rc=$?; tail -2 gpurun_out/bench_f.err; cat gpurun_out/bench_f.json
