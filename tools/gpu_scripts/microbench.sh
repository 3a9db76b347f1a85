#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/microbench.py > gpurun_out/microbench.log 2>&1; echo "rc=$?"
timeout 600 python tools/microbench_warps.py > gpurun_out/microbench_warps.log 2>&1; echo "rc=$?"
tail -30 gpurun_out/microbench_warps.log
