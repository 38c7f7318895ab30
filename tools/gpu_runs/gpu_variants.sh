for v in "-DSPRL_EVALNET_UNIT_KSTEPS=1" "-DSPRL_EVALNET_UNIT_KSTEPS=4" "-DSPRL_EVALNET_UNIT_KSTEPS=2 -DSPRL_EVALNET_CLUSTER=1"; do
  echo "=== $v"; tools/build_variant.sh $v || continue
  timeout 120 python tools/check_evalnet.py 257 2 2>&1 | grep "dlogit"
  SPRL_EVALNET_TIMING=1 timeout 120 python tools/check_evalnet.py 32768 2 2>&1 | grep "forward B\|CTA 0" | tail -2
done
