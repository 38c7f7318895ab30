set -x
nvidia-smi -L | head -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_m_8gpu.json 2> gpurun_out/bench_m_8gpu.err; echo rc=$?; tail -3 gpurun_out/bench_m_8gpu.err; cut -c1-700 gpurun_out/bench_m_8gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 3 --warmup 3 --no-e2e > gpurun_out/bench_m_4gpu.json 2> gpurun_out/bench_m_4gpu.err; echo rc=$?; cut -c1-300 gpurun_out/bench_m_4gpu.json
