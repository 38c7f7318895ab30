run() { timeout 100 python tools/check_evalnet.py 257 2 2>&1 | grep "dlogit" | cut -c1-80; for b in 32768 65536; do timeout 100 python tools/check_evalnet.py $b 2 2>&1 | grep "forward B\|rror" | tail -1; done; }
echo "=== ld x3 (default)"; tools/build_variant.sh && run
echo "=== ld split (old)"; tools/build_variant.sh -DSPRL_EVALNET_LD_SPLIT && run
echo "=== ld x3 + MMA order hi*Whi, lo*Whi, hi*Wlo"; tools/build_variant.sh -DSPRL_EVALNET_ORDER_B && run
echo "=== ld x3 (default) again"; tools/build_variant.sh && run
