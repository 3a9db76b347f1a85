#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests/ -x -q -m gpu --durations=8 > gpurun_out/t_full.log 2>&1
echo "full gpu tests rc=$?"; tail -20 gpurun_out/t_full.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
