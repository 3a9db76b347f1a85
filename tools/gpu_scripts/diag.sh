#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "rank_normalised or cuqi or reference_shaped" > gpurun_out/t_diag.log 2>&1; echo rc=$?; tail -25 gpurun_out/t_diag.log
