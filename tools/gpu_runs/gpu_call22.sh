set -x
timeout 900 python -m pytest tests/test_evalnet.py -m gpu -x -q 2>&1 | tail -3
timeout 120 python tools/check_evalnet.py 32768 2 2>&1 | grep "forward B" | tail -1
timeout 120 python tools/check_evalnet.py 65536 2 2>&1 | grep "forward B" | tail -1
tools/build_variant.sh -DSPRL_EVALNET_ROTATE=1
timeout 120 python tools/check_evalnet.py 32768 2 2>&1 | grep "forward B" | tail -1
timeout 120 python tools/check_evalnet.py 65536 2 2>&1 | grep "forward B" | tail -1
