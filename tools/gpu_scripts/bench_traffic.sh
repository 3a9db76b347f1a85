#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 --cpu-seconds 0 > gpurun_out/plain_bench2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tc_run_kernel -s 4 -c 1 -o gpurun_out/prof_bench_r9 python bench.py --steps 3 --warmup 3 --cpu-seconds 0 > gpurun_out/ncu_bench2.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_bench2.log | cut -c1-300
