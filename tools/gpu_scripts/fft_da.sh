#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "fft_coarse or delayed" > gpurun_out/t_fft.log 2>&1
echo "rc=$?"; tail -30 gpurun_out/t_fft.log
