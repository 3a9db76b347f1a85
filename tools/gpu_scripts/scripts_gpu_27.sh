#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench rc=$?"
timeout 900 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
echo "ref rc=$?"
python tools/profile_run.py tc 4096 16 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tc_run_kernel -s 1 -c 1 -o gpurun_out/prof_tc_r7 python tools/profile_run.py tc 4096 16 > gpurun_out/ncu.log 2>&1
echo "ncu rc=$?"
python bench.py --steps 2 --warmup 3 --cpu-seconds 0 > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r7.csv python bench.py --steps 2 --warmup 3 --cpu-seconds 0 > gpurun_out/ncu_bench.log 2>&1
echo "ncu launches rc=$?"
cat gpurun_out/bench_default.json; cat gpurun_out/bench_reference.json
