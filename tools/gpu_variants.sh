for v in "-DSPRL_EVALNET_CLUSTER=4 -DSPRL_EVALNET_TESTWAIT" "-DSPRL_EVALNET_CLUSTER=2 -DSPRL_EVALNET_TESTWAIT" "-DSPRL_EVALNET_CLUSTER=2" "-DSPRL_EVALNET_CLUSTER=1 -DSPRL_EVALNET_TESTWAIT"; do
  echo "=== $v"; tools/build_variant.sh $v || continue
  for d in 0 3; do SPRL_EVALNET_DEBUG=$d timeout 120 python tools/check_evalnet.py 32768 2 2>&1 | grep "forward B"; done
done
