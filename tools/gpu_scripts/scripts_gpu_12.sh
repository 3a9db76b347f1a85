#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/microbench.py > gpurun_out/microbench2.log 2>&1
echo "mb rc=$?"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t_all2.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/t_all2.log
cat gpurun_out/microbench2.log
