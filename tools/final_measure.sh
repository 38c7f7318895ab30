#!/bin/bash
# tools/final_measure.sh <tag> -- the round's closing measurement on one B200 box (run through tools/gpu.sh): the GPU test
# suite, smoke, the bench line, the reference arm, the ncu launch list and one full capture of the step's kernels (each
# ncu pass only after the same command has exited 0 without ncu), and the other configurations.  Outputs: gpurun_out/<tag>_*.
T=${1:-final}
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/${T}_pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -2 $O/${T}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; echo "smoke rc $?"
python bench.py --steps 20 --warmup 5 > $O/${T}_bench_1gpu.json 2> $O/${T}_bench_1gpu.err; echo "bench rc $?"
python bench.py --impl reference --steps 1 --warmup 0 > $O/${T}_bench_reference_arm.json 2> $O/${T}_ref.err; echo "ref rc $?"
python bench.py --steps 2 --warmup 1 --rounds 16 --no-e2e --no-cpu-baseline > $O/${T}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/${T}_launches_bench.csv python bench.py --steps 2 --warmup 1 --rounds 16 --no-e2e --no-cpu-baseline > $O/${T}_ncu1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_evalnet_resident|k_round|k_heads" -s 100 -c 5 -o $O/${T}_bench_kernels python bench.py --steps 2 --warmup 1 --rounds 16 --no-e2e --no-cpu-baseline > $O/${T}_ncu2.log 2>&1
echo "ncu rc $?"
python tools/bench_configs.py > $O/${T}_configs.jsonl 2> $O/${T}_configs.err; echo "configs rc $?"
