#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/microbench.py > gpurun_out/microbench.log 2>&1
echo "rc=$?"; cat gpurun_out/microbench.log
