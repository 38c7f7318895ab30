set -x
python tools/profile_search.py 16384 6 ext 1 1500 > gpurun_out/plain_prof_r1q.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_round -s 1502 -c 2 -o gpurun_out/prof_round_r1q python tools/profile_search.py 16384 6 ext 1 1500 > gpurun_out/ncu_prof_r1q.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
