#!/bin/bash
for K in 0 2 0 2; do
  export LUMO_GPU_SO=$PWD/lumo_b200/liblumo_gpu_k$K.so
  echo -n "K=$K "; python tools/micro_run.py bunny 2>&1 | tail -1
  for s in "bunny 4" "bistro 1"; do
    echo -n "K=$K "; timeout 300 python tools/prof_run.py $s 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['workload'], round(d['ms'],1), {k:round(v[0],1) for k,v in d['kernel_ms'].items()})"
  done
done
