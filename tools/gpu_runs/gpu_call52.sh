timeout 600 python -m pytest tests/test_evalnet.py -m gpu -x -q 2>&1 | tail -2
for b in 32768 65536; do timeout 100 python tools/check_evalnet.py $b 2 2>&1 | grep "forward B\|rror" | tail -1; done
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_chk2.csv python tools/check_evalnet.py 65536 2 > /dev/null 2>&1; python tools/ncu_summary.py launches gpurun_out/launches_chk2.csv | head -5
