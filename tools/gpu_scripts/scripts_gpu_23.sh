#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/microbench.py 15,29 > gpurun_out/microbench11.log 2>&1
MEMS_LIB=$PWD/mcmc-mems_b200/libmems_b200_mb576.so timeout 300 python tools/microbench.py 15,29 > gpurun_out/microbench11b.log 2>&1
echo "mb rc=$?"; grep "^mode" gpurun_out/microbench11.log; echo "--- 112 regs"; grep "^mode" gpurun_out/microbench11b.log
