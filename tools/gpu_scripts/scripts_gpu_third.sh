#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 300 python tools/probe_numerics.py > gpurun_out/probe.log 2>&1
echo "probe rc=$?" >> gpurun_out/summary.txt
timeout 1200 python -m pytest tests/test_gpu.py -q --timeout 300 > gpurun_out/t_all.log 2>&1
echo "pytest rc=$?" >> gpurun_out/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/summary.txt
timeout 300 python bench.py --steps 5 --warmup 3 --cpu-seconds 5 > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err
echo "bench tc rc=$?" >> gpurun_out/summary.txt
timeout 300 python bench.py --steps 5 --warmup 3 --cpu-seconds 0 --chains 18944 --inner 64 --thin 32 > gpurun_out/bench_tc_18944.json 2> gpurun_out/bench_tc2.err
echo "bench tc 18944 rc=$?" >> gpurun_out/summary.txt
python tools/profile_run.py tc 4096 16 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tc_run_kernel -s 1 -c 1 -o gpurun_out/prof_tc_r1 python tools/profile_run.py tc 4096 16 > gpurun_out/ncu.log 2>&1
echo "ncu rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -4 gpurun_out/t_all.log; tail -2 gpurun_out/smoke.log; grep -E "forward|loglik|split" gpurun_out/probe.log
cut -c1-300 gpurun_out/bench_tc.json; echo; cut -c1-300 gpurun_out/bench_tc_18944.json
