#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests/ -x -q -m gpu > gpurun_out/t_full6.log 2>&1
echo "full gpu tests rc=$?"; tail -2 gpurun_out/t_full6.log
for i in 1 2; do timeout 300 python bench.py --steps 5 --warmup 3 --cpu-seconds 0 2>/dev/null | tail -1 | cut -c1-100; done
