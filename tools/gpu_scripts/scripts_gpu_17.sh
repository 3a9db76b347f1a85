#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/microbench.py 20,21,22,23,24,25,26,27 > gpurun_out/microbench5.log 2>&1
echo "mb rc=$?"; cat gpurun_out/microbench5.log
