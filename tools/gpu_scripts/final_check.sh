#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests/ -x -q -m gpu > gpurun_out/t_final.log 2>&1
echo "full gpu tests rc=$?"; tail -2 gpurun_out/t_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2.err; echo "bench rc=$?"; wc -l gpurun_out/bench_final2.json; cut -c1-150 gpurun_out/bench_final2.json
timeout 900 python bench.py --impl reference > gpurun_out/bench_final2_ref.json 2>/dev/null; echo "ref rc=$?"; wc -l gpurun_out/bench_final2_ref.json; cut -c1-150 gpurun_out/bench_final2_ref.json
