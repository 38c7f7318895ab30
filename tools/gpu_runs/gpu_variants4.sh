run() { for b in 32768 65536; do timeout 100 python tools/check_evalnet.py $b 2 2>&1 | grep "forward B\|rror" | tail -1; done; }
echo "=== default (cluster 2)"; tools/build_variant.sh && run
echo "=== cluster 1"; tools/build_variant.sh -DSPRL_EVALNET_CLUSTER=1 && run
echo "=== cluster 2 again"; tools/build_variant.sh && run
echo "=== cluster 1 again"; tools/build_variant.sh -DSPRL_EVALNET_CLUSTER=1 && run
tools/build_variant.sh
