#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
echo "n8 rc=$?"; tail -3 gpurun_out/bench_n8.err; cut -c1-250 gpurun_out/bench_n8.json
