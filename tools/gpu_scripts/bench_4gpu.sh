#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err
echo "n4 rc=$?"; cut -c1-160 gpurun_out/bench_n4.json
