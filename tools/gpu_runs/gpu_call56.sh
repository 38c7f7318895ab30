# k_perft_expand (warp-cooperative, one atomic per warp, frontier arena) and k_emit (one thread per (symmetry, cell)): parity + ncu
set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/profile_env.py 11 22 2048 2>&1 | tail -8
timeout 300 python tools/profile_env.py 10 22 2048 2>&1 | head -1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_perft|k_emit' -c 24 \
  -o gpurun_out/prof_env_r1s -f python tools/profile_env.py 10 16 2048 > gpurun_out/ncu_env_r1s.log 2>&1; echo ncu env rc=$?
