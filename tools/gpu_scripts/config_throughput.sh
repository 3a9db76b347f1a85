#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/config_throughput.py > gpurun_out/config_throughput.log 2>&1; echo "rc=$?"; cat gpurun_out/config_throughput.log | tail -8
timeout 300 python -m pytest tests -m gpu -q -x -k "reference_shaped_call" 2>&1 | tail -2
