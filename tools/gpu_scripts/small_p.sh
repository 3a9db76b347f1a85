#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "one_and_two_parameter" > gpurun_out/t_smallp.log 2>&1; echo rc=$?; tail -25 gpurun_out/t_smallp.log
