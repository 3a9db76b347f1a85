#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 1500 python -m pytest tests/test_gpu.py -q --timeout 400 > gpurun_out/t_all.log 2>&1
echo "pytest rc=$?" >> gpurun_out/summary.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "bench n2 rc=$?" >> gpurun_out/summary.txt
timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --cpu-seconds 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench n1 rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; grep -E "^E  |^FAILED|passed|failed" gpurun_out/t_all.log | head -20
cut -c1-140 gpurun_out/bench_n2.json; echo; cut -c1-140 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n2.err
