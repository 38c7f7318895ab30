# end-of-round verification after the k_round / k_emit / perft changes + ncu launch list of the bench command
set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_w.json 2> gpurun_out/bench_w.err; echo bench rc=$?
cat gpurun_out/bench_w.json
python bench.py --steps 1 --warmup 3 --rounds 16 --no-e2e --no-cpu-baseline > gpurun_out/plain_bench_r1w.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 200 --csv --log-file gpurun_out/launches_r1w.csv python bench.py --steps 1 --warmup 3 --rounds 16 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches_r1w.log 2>&1; echo launches rc=$?
timeout 300 python tools/profile_env.py 11 22 2048 2>&1 | head -2
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_w_ref.json 2> gpurun_out/bench_w_ref.err; cat gpurun_out/bench_w_ref.json
