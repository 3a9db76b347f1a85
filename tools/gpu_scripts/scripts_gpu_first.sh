#!/bin/bash
# first GPU contact: plumbing self-test, both paths' parity tests in separate processes, then bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 300 python -m pytest tests/test_gpu.py -q --timeout 120 -k "umma_selftest" > gpurun_out/t_selftest.log 2>&1
echo "selftest rc=$?" >> gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu.py -q --timeout 300 -k "fp32 or philox_stream or philox_proposals or predict" > gpurun_out/t_fp32.log 2>&1
echo "fp32 rc=$?" >> gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu.py -q --timeout 300 -k "tc and not umma" > gpurun_out/t_tc.log 2>&1
echo "tc rc=$?" >> gpurun_out/summary.txt
timeout 300 python -m pytest tests/test_gpu.py -q --timeout 200 -k "end_to_end" > gpurun_out/t_api.log 2>&1
echo "api rc=$?" >> gpurun_out/summary.txt
timeout 300 python bench.py --path fp32 --steps 3 --warmup 3 --inner 64 --thin 32 --cpu-seconds 0 > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err
echo "bench fp32 rc=$?" >> gpurun_out/summary.txt
timeout 300 python bench.py --steps 5 --warmup 3 --cpu-seconds 5 > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err
echo "bench tc rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -5 gpurun_out/t_selftest.log gpurun_out/t_fp32.log gpurun_out/t_tc.log gpurun_out/t_api.log
cat gpurun_out/bench_fp32.json gpurun_out/bench_tc.json | cut -c1-600
