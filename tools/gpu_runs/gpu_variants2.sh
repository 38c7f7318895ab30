# fp16-split evaluator: accuracy and time of the build variants (ring unit size, CTAs per SM, cluster size)
for v in "-DSPRL_EVALNET_UNIT_KSTEPS=4" "-DSPRL_EVALNET_UNIT_KSTEPS=2 -DSPRL_EVALNET_CTAS_PER_SM=2" "-DSPRL_EVALNET_UNIT_KSTEPS=1 -DSPRL_EVALNET_CTAS_PER_SM=2" "-DSPRL_EVALNET_UNIT_KSTEPS=2 -DSPRL_EVALNET_CTAS_PER_SM=2 -DSPRL_EVALNET_CLUSTER=1" "-DSPRL_EVALNET_UNIT_KSTEPS=2" "-DSPRL_EVALNET_UNIT_KSTEPS=1 -DSPRL_EVALNET_CTAS_PER_SM=2 -DSPRL_EVALNET_CLUSTER=4"; do
  echo "=== $v"; tools/build_variant.sh $v || continue
  timeout 120 python tools/check_evalnet.py 257 2 2>&1 | grep "dlogit\|Error\|error"
  SPRL_EVALNET_TIMING=1 timeout 120 python tools/check_evalnet.py 32768 2 2>&1 | grep "forward B\|CTA 0\|Error\|error" | tail -3
done
tools/build_variant.sh -DSPRL_EVALNET_UNIT_KSTEPS=4
timeout 600 python -m pytest tests/test_evalnet.py -m gpu -x -q 2>&1 | tail -15
