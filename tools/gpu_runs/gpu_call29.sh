for d in 0 6; do
echo "=== SPRL_EVALNET_DEBUG=$d"
SPRL_EVALNET_DEBUG=$d SPRL_EVALNET_TIMING=1 timeout 120 python tools/check_evalnet.py 32768 2 2>&1 | grep "forward B\|CTA 0" | tail -2
done
