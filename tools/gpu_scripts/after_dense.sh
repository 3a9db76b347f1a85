#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests/ -x -q -m gpu > gpurun_out/t_full5.log 2>&1
echo "full gpu tests rc=$?"; tail -3 gpurun_out/t_full5.log
timeout 900 python tools/config_throughput.py 2>&1 | tail -6 | tee gpurun_out/config_throughput2.log
timeout 300 python bench.py --steps 5 --warmup 3 --cpu-seconds 0 2>/dev/null | tail -1 | cut -c1-120
