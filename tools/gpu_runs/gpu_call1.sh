set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
bash tools/worker_smoke.sh > gpurun_out/worker_smoke.log 2>&1; echo worker rc=$?

python tools/explore_nn.py 32768 > gpurun_out/explore_nn.log 2>&1
python bench.py > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; echo bench rc=$?
python bench.py --steps 1 --warmup 3 --rounds 16 --no-e2e --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 3 --rounds 16 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
python tools/profile_search.py 8192 6 ext 1 1500 > gpurun_out/plain_prof.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_round -s 1502 -c 3 -o gpurun_out/prof_round_r1b python tools/profile_search.py 8192 6 ext 1 1500 > gpurun_out/ncu_prof.log 2>&1
cat gpurun_out/plain_prof.log
