run() { timeout 100 python tools/check_evalnet.py 257 2 2>&1 | grep "dlogit\|rror" | cut -c1-70; for b in 32768 65536; do timeout 100 python tools/check_evalnet.py $b 2 2>&1 | grep "forward B\|rror" | tail -1; done; }
echo "=== default"; tools/build_variant.sh && run
for h in 1000 20000 1000000; do echo "=== wait hint $h ns"; tools/build_variant.sh -DSPRL_EVALNET_WAIT_HINT_NS=$h && run; done
echo "=== default again"; tools/build_variant.sh && run
