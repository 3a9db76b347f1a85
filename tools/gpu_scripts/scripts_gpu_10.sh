#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
run() { # name lib nh
  for cfg in "4096 256 64" "18944 64 32"; do set -- $1 $2 $3 $cfg
    MEMS_LIB=$2 MEMS_TC_NH=$3 timeout 300 python bench.py --steps 5 --warmup 3 --cpu-seconds 0 --chains $4 --inner $5 --thin $6 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 nh$3 chains $4:', round(d['value']/1e6,2), 'M steps/s')" >> gpurun_out/summary.txt
  done
}
L=$PWD/mcmc-mems_b200
run base $L/libmems_b200.so 2
run hint $L/libmems_b200_hint.so 2
run g4 $L/libmems_b200_g4.so 2
run base $L/libmems_b200.so 1
run g4 $L/libmems_b200_g4.so 1
cat gpurun_out/summary.txt
