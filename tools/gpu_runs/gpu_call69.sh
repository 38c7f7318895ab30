# last verification of the round: GPU tests, smoke, bench, reference arm
set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err; echo bench rc=$?
cat gpurun_out/bench_x.json
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_x_ref.json 2> gpurun_out/bench_x_ref.err; cat gpurun_out/bench_x_ref.json
