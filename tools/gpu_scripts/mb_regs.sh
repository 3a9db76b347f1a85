#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/microbench.py 11 2>&1 | grep "^mode" 
echo "--- 96-register budget"
MEMS_LIB=$PWD/mcmc-mems_b200/libmems_b200_mb576.so timeout 300 python tools/microbench.py 11 2>&1 | grep "^mode"
