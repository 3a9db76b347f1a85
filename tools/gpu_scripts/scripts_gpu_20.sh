#!/bin/bash
mkdir -p gpurun_out
MB_STAGGER=1 timeout 300 python tools/microbench.py 11,15,17 > gpurun_out/microbench8.log 2>&1
echo "mb rc=$?"; grep "^mode" gpurun_out/microbench8.log
