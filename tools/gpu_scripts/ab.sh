#!/bin/bash
# A/B of library variants on the bench workload: bash tools/gpu_scripts/ab.sh out.log default name=DEF ...  (variants built beforehand)
out=$1; shift
mkdir -p gpurun_out
timeout 1200 python tools/ab_bench.py "$@" > gpurun_out/$out 2>&1; echo "ab rc=$?"; cat gpurun_out/$out
