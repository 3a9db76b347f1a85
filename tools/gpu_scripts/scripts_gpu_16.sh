#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/microbench.py 3,11,12,13,14 > gpurun_out/microbench4.log 2>&1
echo "mb rc=$?"; grep -v "^mma" gpurun_out/microbench4.log
