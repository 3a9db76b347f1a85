#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary10.txt
run() { # name lib
  for cfg in "4096 256 64" "18944 64 32"; do set -- $1 $2 $cfg
    MEMS_LIB=$2 timeout 300 python bench.py --steps 5 --warmup 3 --cpu-seconds 0 --chains $3 --inner $4 --thin $5 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 chains $3:', round(d['value']/1e6,2), 'M steps/s')" >> gpurun_out/summary10.txt
  done
}
L=$PWD/mcmc-mems_b200
run base $L/libmems_b200.so
run r88 $L/libmems_b200_r88.so
run r80 $L/libmems_b200_r80.so
cat gpurun_out/summary10.txt
