set -x
timeout 900 python -m pytest tests/test_evalnet.py -m gpu -x -q 2>&1 | tail -3
SPRL_EVALNET_TIMING=1 timeout 120 python tools/check_evalnet.py 32768 2 2>&1 | grep "forward B\|dlogit\|CTA 0" | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_chk.csv python tools/check_evalnet.py 32768 2 > /dev/null 2>&1; python tools/ncu_summary.py launches gpurun_out/launches_chk.csv | head -5
