set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err; echo bench rc=$?; tail -2 gpurun_out/bench_d.err; cat gpurun_out/bench_d.json
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_d_ref.json 2> gpurun_out/bench_d_ref.err; cat gpurun_out/bench_d_ref.json
bash tools/worker_smoke.sh 2>&1 | tail -3
