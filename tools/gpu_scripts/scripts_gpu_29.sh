#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary6.txt
timeout 900 python -m pytest tests -m gpu -q -x -k "tc" > gpurun_out/t_q.log 2>&1
echo "tests rc=$?"; tail -12 gpurun_out/t_q.log
run() { # name q
  for cfg in "4096 256 64" "18944 64 32"; do set -- $1 $2 $cfg
    MEMS_TC_Q=$2 timeout 300 python bench.py --steps 5 --warmup 3 --cpu-seconds 0 --chains $3 --inner $4 --thin $5 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 chains $3:', round(d['value']/1e6,2), 'M steps/s', d['check']['post_std'])" >> gpurun_out/summary6.txt
  done
}
run tcq 1
run old 0
cat gpurun_out/summary6.txt
