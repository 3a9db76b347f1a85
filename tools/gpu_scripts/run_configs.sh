#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_configs.py -m gpu -q -x --durations=10 > gpurun_out/t_cfg.log 2>&1
echo "cfg tests rc=$?"; tail -40 gpurun_out/t_cfg.log
