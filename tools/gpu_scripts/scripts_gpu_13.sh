#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary2.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/t_all3.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/t_all3.log
run() { # name lib
  for cfg in "4096 256 64" "18944 64 32"; do set -- $1 $2 $cfg
    MEMS_LIB=$2 timeout 300 python bench.py --steps 5 --warmup 3 --cpu-seconds 0 --chains $3 --inner $4 --thin $5 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 chains $3:', round(d['value']/1e6,2), 'M steps/s')" >> gpurun_out/summary2.txt
  done
}
L=$PWD/mcmc-mems_b200
run base $L/libmems_b200.so
run exp1 $L/libmems_b200_exp1.so
run exp2 $L/libmems_b200_exp2.so
cat gpurun_out/summary2.txt
