# other configurations with the round's last build
timeout 420 python tools/bench_configs.py > gpurun_out/bench_other_x.jsonl 2> gpurun_out/bench_other_x.err; echo rc=$?
grep -c config gpurun_out/bench_other_x.jsonl
