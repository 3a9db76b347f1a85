#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "edge_sizes" > gpurun_out/t_edge.log 2>&1; echo rc=$?; tail -25 gpurun_out/t_edge.log
