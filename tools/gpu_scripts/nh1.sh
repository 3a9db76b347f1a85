#!/bin/bash
mkdir -p gpurun_out
MEMS_TC_NH=1 timeout 1200 python -m pytest tests -m gpu -q -x -k "tc or config or coarse or delayed" > gpurun_out/t_nh1.log 2>&1; echo rc=$?; tail -5 gpurun_out/t_nh1.log
MEMS_TC_NH=1 timeout 300 python bench.py --steps 5 --warmup 3 --cpu-seconds 0 2>/dev/null | tail -1 | cut -c1-120
