#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/microbench.py > gpurun_out/microbench3.log 2>&1
echo "mb rc=$?"; tail -8 gpurun_out/microbench3.log
