set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "match_external" 2>&1 | tail -15
timeout 600 bash tools/evaluate_smoke.sh 2>&1 | tail -25
timeout 300 bash tools/worker_smoke.sh 2>&1 | tail -4
