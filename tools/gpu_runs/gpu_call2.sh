set -x
timeout 120 python tools/check_evalnet.py 2 2 2>&1 | tail -12
timeout 120 python tools/check_evalnet.py 257 2 2>&1 | tail -12
timeout 120 python tools/check_evalnet.py 32768 2 2>&1 | tail -12
timeout 120 python tools/check_evalnet.py 301 0 2>&1 | tail -6
timeout 120 python tools/check_evalnet.py 301 6 2>&1 | tail -6
timeout 600 python -m pytest tests -m gpu -x -q -k "external or sharded or traced" 2>&1 | tail -8
