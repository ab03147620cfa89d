#!/bin/bash
# compare library build variants: tools/gpu_variants.sh <suffixes...>  (lumo_b200/liblumo_gpu_<suffix>.so)
for V in "$@"; do
  export LUMO_GPU_SO=$PWD/lumo_b200/liblumo_gpu_$V.so
  for s in "bunny 4" "bistro 1" "conference 4" "caustics_bdpt 1"; do
    echo -n "$V "; timeout 300 python tools/prof_run.py $s 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['workload'], round(d['ms'],1), {k:round(v[0],1) for k,v in d['kernel_ms'].items()})"
  done
done
