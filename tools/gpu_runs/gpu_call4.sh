set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
bash tools/worker_smoke.sh > gpurun_out/worker_smoke.log 2>&1; echo worker rc=$?; tail -4 gpurun_out/worker_smoke.log
bash tools/worker_smoke.sh sprl_b200/host/bin/ref_OTHWorker > gpurun_out/worker_smoke_ref.log 2>&1; echo refworker rc=$?; tail -3 gpurun_out/worker_smoke_ref.log
python bench.py > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err; echo bench rc=$?; tail -3 gpurun_out/bench_c.err; cat gpurun_out/bench_c.json
python bench.py --evaluator libtorch --no-e2e --no-cpu-baseline > gpurun_out/bench_c_libtorch.json 2>gpurun_out/bench_c_libtorch.err; cat gpurun_out/bench_c_libtorch.json
