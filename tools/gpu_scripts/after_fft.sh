#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests/ -x -q -m gpu > gpurun_out/t_full4.log 2>&1
echo "full gpu tests rc=$?"; tail -3 gpurun_out/t_full4.log
timeout 600 python tools/da_timing.py 262144 40 2>&1 | tail -8 | tee gpurun_out/da_timing.log
