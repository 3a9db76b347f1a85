#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 1200 python -m pytest tests/test_gpu.py -q --timeout 300 -k "tc" > gpurun_out/t_tc_nh2.log 2>&1
echo "pytest nh2 rc=$?" >> gpurun_out/summary.txt
for nh in 1 2; do
MEMS_TC_NH=$nh timeout 300 python bench.py --steps 5 --warmup 3 --cpu-seconds 0 > gpurun_out/bench_tc_nh$nh.json 2> gpurun_out/bench_tc.err
echo "bench tc nh$nh rc=$?" >> gpurun_out/summary.txt
MEMS_TC_NH=$nh timeout 300 python bench.py --steps 5 --warmup 3 --cpu-seconds 0 --chains 18944 --inner 64 --thin 32 > gpurun_out/bench_tc_18944_nh$nh.json 2> gpurun_out/bench_tc2.err
echo "bench tc 18944 nh$nh rc=$?" >> gpurun_out/summary.txt
done
python tools/profile_run.py tc 4096 16 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tc_run_kernel -s 1 -c 1 -o gpurun_out/prof_tc_r4 python tools/profile_run.py tc 4096 16 > gpurun_out/ncu.log 2>&1
echo "ncu rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -3 gpurun_out/t_tc_nh2.log
for f in gpurun_out/bench_tc_nh1.json gpurun_out/bench_tc_nh2.json gpurun_out/bench_tc_18944_nh1.json gpurun_out/bench_tc_18944_nh2.json; do cut -c1-110 $f; echo; done
