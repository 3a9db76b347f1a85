#!/bin/bash
mkdir -p gpurun_out
python tools/profile_run.py tc 4096 16 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tcq_run_kernel -s 1 -c 1 -o gpurun_out/prof_tcq_r8 python tools/profile_run.py tc 4096 16 > gpurun_out/ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu.log
