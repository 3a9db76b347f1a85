#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2b.json 2> gpurun_out/bench_n2b.err
echo "n2 rc=$?"; tail -3 gpurun_out/bench_n2b.err; cut -c1-400 gpurun_out/bench_n2b.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err
echo "ref n2 rc=$?"; cut -c1-300 gpurun_out/bench_ref_n2.json
