set -x
# launch list of the benchmark command (shares, not absolutes)
python bench.py --steps 1 --warmup 3 --rounds 16 --no-e2e --no-cpu-baseline > gpurun_out/plain_bench_r1p.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 200 --csv --log-file gpurun_out/launches_r1p.csv python bench.py --steps 1 --warmup 3 --rounds 16 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches_r1p.log 2>&1
# full captures: the evaluator on the benchmark's leaf batch, and the search launch in mid-game, at the bench's 16384 games
python tools/check_evalnet.py 65536 2 > gpurun_out/evalnet_plain_r1p.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_evalnet -s 2 -c 1 -o gpurun_out/prof_evalnet_r1p python tools/check_evalnet.py 65536 2 > gpurun_out/ncu_evalnet_r1p.log 2>&1
python tools/profile_search.py 16384 6 ext 1 1500 > gpurun_out/plain_prof_r1p.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_round -s 1502 -c 2 -o gpurun_out/prof_round_r1p python tools/profile_search.py 16384 6 ext 1 1500 > gpurun_out/ncu_prof_r1p.log 2>&1
tail -n 3 gpurun_out/plain_prof_r1p.log gpurun_out/evalnet_plain_r1p.log
