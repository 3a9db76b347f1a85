#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/microbench.py 15,28 > gpurun_out/microbench10.log 2>&1
echo "mb rc=$?"; grep "^mode" gpurun_out/microbench10.log
