set -x
timeout 120 python tools/check_evalnet.py 257 2 2>&1 | tail -5
timeout 120 python tools/check_evalnet.py 301 6 2>&1 | tail -4
timeout 120 python tools/check_evalnet.py 32768 2 > gpurun_out/evalnet_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_evalnet -s 2 -c 1 -o gpurun_out/prof_evalnet_v2 python tools/check_evalnet.py 32768 2 > gpurun_out/ncu_evalnet.log 2>&1
cat gpurun_out/evalnet_plain.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
