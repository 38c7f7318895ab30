# end-of-round verification + ncu captures of the kernels outside the search round
set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_r.json 2> gpurun_out/bench_r.err; echo bench rc=$?
cat gpurun_out/bench_r.json
timeout 300 python tools/profile_env.py 10 22 2048 2>&1 | tail -8
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_rollout|k_perft|k_env_step|k_emit|k_begin' -c 40 \
  -o gpurun_out/prof_env_r1r -f python tools/profile_env.py 10 22 2048 > gpurun_out/ncu_env_r1r.log 2>&1; echo ncu env rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_heads' -c 2 \
  -o gpurun_out/prof_heads_r1r -f python tools/check_evalnet.py 65536 2 > gpurun_out/ncu_heads_r1r.log 2>&1; echo ncu heads rc=$?
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_r_ref.json 2> gpurun_out/bench_r_ref.err; cat gpurun_out/bench_r_ref.json
