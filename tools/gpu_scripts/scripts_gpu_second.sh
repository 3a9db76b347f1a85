#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 300 python tools/probe_numerics.py > gpurun_out/probe.log 2>&1
echo "probe rc=$?" >> gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu.py -q --timeout 300 -k "fp32 or philox_stream or philox_proposals or predict" > gpurun_out/t_fp32.log 2>&1
echo "fp32 rc=$?" >> gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu.py -q --timeout 300 -k "tc and not umma" > gpurun_out/t_tc.log 2>&1
echo "tc rc=$?" >> gpurun_out/summary.txt
timeout 300 python -m pytest tests/test_gpu.py -q --timeout 200 -k "end_to_end" > gpurun_out/t_api.log 2>&1
echo "api rc=$?" >> gpurun_out/summary.txt
timeout 300 python bench.py --path fp32 --steps 3 --warmup 3 --inner 64 --thin 32 --cpu-seconds 0 > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err
echo "bench fp32 rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; cat gpurun_out/probe.log
