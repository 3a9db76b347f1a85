#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "autocov or least_squares or sensitivity" > gpurun_out/t_new.log 2>&1
echo "new tests rc=$?"; tail -40 gpurun_out/t_new.log
