set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py > gpurun_out/bench_k.json 2> gpurun_out/bench_k.err; echo bench rc=$?; tail -2 gpurun_out/bench_k.err; cat gpurun_out/bench_k.json
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_k_ref.json 2> gpurun_out/bench_k_ref.err; cat gpurun_out/bench_k_ref.json
