#!/bin/bash
# quick perf check: bunny + bistro + conference short renders, then the GPU test-suite
mkdir -p gpurun_out
for s in "bunny 4" "bistro 1" "conference 4"; do
  python tools/prof_run.py $s 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['workload'], round(d['ms'],1), {k:round(v[0],1) for k,v in d['kernel_ms'].items()})"
done
make -s -C oracle liblumo_oracle.so 2>/dev/null
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
