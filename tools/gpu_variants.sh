for v in "-DSPRL_EVALNET_UNIT_KSTEPS=8" "-DSPRL_EVALNET_UNIT_KSTEPS=4"; do
  echo "=== $v"; tools/build_variant.sh $v || continue
  timeout 120 python tools/check_evalnet.py 257 2 2>&1 | grep "dlogit"
  for d in 0 3; do SPRL_EVALNET_TIMING=1 SPRL_EVALNET_DEBUG=$d timeout 120 python tools/check_evalnet.py 32768 2 2>&1 | grep "forward B\|CTA 0" | tail -2; done
done
