timeout 100 python tools/check_evalnet.py 257 2 2>&1 | grep "dlogit\|rror\|Traceback" | head -3
SPRL_EVALNET_TIMING=1 timeout 100 python tools/check_evalnet.py 32768 2 2>&1 | grep "forward B\|CTA 0\|rror" | tail -2
timeout 100 python tools/check_evalnet.py 65536 2 2>&1 | grep "forward B\|rror" | tail -1
timeout 600 python -m pytest tests/test_evalnet.py -m gpu -x -q 2>&1 | tail -4
