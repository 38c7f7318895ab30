set -x
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "drop or heuristic or match" 2>&1 | tail -30
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
