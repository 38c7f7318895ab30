# 2 GPUs with the round's final kernels
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/bench_2gpu_w.json 2> gpurun_out/bench_2gpu_w.err; echo rc=$?
cat gpurun_out/bench_2gpu_w.json | cut -c1-300
