#!/bin/bash
# kd inner-loop bound sweep (LUMO_KD_ROUND): nested default and node-major
for K in 1000000 2 4 8; do
  export LUMO_GPU_SO=$PWD/lumo_b200/liblumo_gpu_k$K.so
  for s in "bunny 4" "bistro 1" "conference 4"; do
    echo "K=$K NM=0 $s"; LUMO_TRACE_NM=0 timeout 300 python tools/prof_run.py $s 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms'],1), {k:round(v[0],1) for k,v in d['kernel_ms'].items()})"
  done
  echo "K=$K NM=1 bunny 4 cold"; LUMO_TRACE_NM=1 PROF_WARM=0 timeout 300 python tools/prof_run.py bunny 4 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms'],1), {k:round(v[0],1) for k,v in d['kernel_ms'].items()})"
done
