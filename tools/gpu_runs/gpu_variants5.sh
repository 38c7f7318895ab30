run() { timeout 100 python tools/check_evalnet.py 257 2 2>&1 | grep "dlogit" | cut -c1-80; for b in 32768 65536; do timeout 100 python tools/check_evalnet.py $b 2 2>&1 | grep "forward B\|rror" | tail -1; done; }
echo "=== prefetch"; tools/build_variant.sh && run
echo "=== no prefetch"; tools/build_variant.sh -DSPRL_EVALNET_NO_PREFETCH && run
echo "=== prefetch again"; tools/build_variant.sh && run
timeout 600 python -m pytest tests/test_evalnet.py -m gpu -x -q 2>&1 | tail -2
