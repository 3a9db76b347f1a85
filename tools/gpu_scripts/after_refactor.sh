#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests/ -x -q -m gpu > gpurun_out/t_full2.log 2>&1
echo "full gpu tests rc=$?"; tail -3 gpurun_out/t_full2.log
timeout 300 python tools/microbench.py 11,15,20 > gpurun_out/microbench12.log 2>&1; echo "mb rc=$?"; grep "warps 16" gpurun_out/microbench12.log
timeout 300 python bench.py --steps 5 --warmup 3 --cpu-seconds 0 2>/dev/null | tail -1 | cut -c1-200
