#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 1500 python -m pytest tests/test_gpu.py -q --timeout 400 > gpurun_out/t_all.log 2>&1
echo "pytest rc=$?" >> gpurun_out/summary.txt
timeout 300 python bench.py --steps 5 --warmup 3 --cpu-seconds 0 > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err
echo "bench tc rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; grep -E "^E  |^FAILED|passed|failed" gpurun_out/t_all.log | head -30
cut -c1-110 gpurun_out/bench_tc.json
