#!/bin/bash
# bench numbers first (never under a profiler), then the ncu launch list and one full capture of the same command
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
python bench.py --steps 2 --warmup 3 --cpu-seconds 0 > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --cpu-seconds 0 > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:tc_run_kernel -s 4 -c 1 -o gpurun_out/prof_bench python bench.py --steps 3 --warmup 3 --cpu-seconds 0 > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
cut -c1-300 gpurun_out/bench_n1.json
