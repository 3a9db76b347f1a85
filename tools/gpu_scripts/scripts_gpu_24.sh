#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary4.txt
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/t_all5.log 2>&1
echo "tests rc=$?"; tail -15 gpurun_out/t_all5.log
run() { # name pipe
  for cfg in "4096 256 64" "18944 64 32"; do set -- $1 $2 $cfg
    MEMS_TC_PIPE=$2 timeout 300 python bench.py --steps 5 --warmup 3 --cpu-seconds 0 --chains $3 --inner $4 --thin $5 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 chains $3:', round(d['value']/1e6,2), 'M steps/s')" >> gpurun_out/summary4.txt
  done
}
run pipe 1
run old 0
cat gpurun_out/summary4.txt
