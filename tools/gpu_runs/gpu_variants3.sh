# cta_group::2 (pair) evaluator variants; every python call under a timeout so that a protocol bug cannot hold the GPU
run() {
  timeout 100 python tools/check_evalnet.py 257 2 2>&1 | grep "dlogit\|rror\|Traceback" | head -3
  SPRL_EVALNET_TIMING=1 timeout 100 python tools/check_evalnet.py 32768 2 2>&1 | grep "forward B\|CTA 0\|rror" | tail -2
}
echo "=== pair, 12 KB slots, nst auto"; tools/build_variant.sh -DSPRL_EVALNET_PAIR=1 && run
for n in 2 3 4; do echo "=== pair, 12 KB slots, nst $n"; SPRL_EVALNET_NST=$n run; done
echo "=== pair, 24 KB slots (UNIT_KS=4)"; tools/build_variant.sh -DSPRL_EVALNET_PAIR=1 -DSPRL_EVALNET_UNIT_KSTEPS=4 && run
echo "=== pair, 24 KB slots, nst 2"; SPRL_EVALNET_NST=2 run
echo "=== no pair (baseline)"; tools/build_variant.sh -DSPRL_EVALNET_PAIR=0 && run
