#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 1500 python -m pytest tests/test_gpu.py -q --timeout 400 > gpurun_out/t_all.log 2>&1
echo "pytest rc=$?" >> gpurun_out/summary.txt
timeout 300 python tools/probe_numerics.py > gpurun_out/probe.log 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 --cpu-seconds 0 > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err
echo "bench tc rc=$?" >> gpurun_out/summary.txt
timeout 300 python bench.py --steps 5 --warmup 3 --cpu-seconds 0 --chains 18944 --inner 64 --thin 32 > gpurun_out/bench_tc_18944.json 2> gpurun_out/bench_tc2.err
echo "bench tc 18944 rc=$?" >> gpurun_out/summary.txt
timeout 300 python bench.py --steps 3 --warmup 3 --cpu-seconds 0 --chains 151552 --inner 16 --thin 16 > gpurun_out/bench_tc_151552.json 2> gpurun_out/bench_tc3.err
echo "bench tc 151552 rc=$?" >> gpurun_out/summary.txt
python tools/profile_run.py tc 4096 16 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tc_run_kernel -s 1 -c 1 -o gpurun_out/prof_tc_r5 python tools/profile_run.py tc 4096 16 > gpurun_out/ncu.log 2>&1
echo "ncu rc=$?" >> gpurun_out/summary.txt
python bench.py --steps 2 --warmup 3 --cpu-seconds 0 > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --cpu-seconds 0 > gpurun_out/ncu_bench.log 2>&1
echo "ncu launches rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; grep -E "^E  |^FAILED|passed|failed" gpurun_out/t_all.log | head -20; grep -E "forward|loglik" gpurun_out/probe.log
for f in gpurun_out/bench_tc.json gpurun_out/bench_tc_18944.json gpurun_out/bench_tc_151552.json; do cut -c1-110 $f; echo; done
