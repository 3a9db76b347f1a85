#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/microbench.py 11,15,17 > gpurun_out/microbench7.log 2>&1
echo "mb rc=$?"; grep "^mode" gpurun_out/microbench7.log
