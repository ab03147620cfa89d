#!/bin/bash
for t in 0 2048 8192 32768 131072; do
  echo -n "tail $t: "; LUMO_BW_TAIL=$t timeout 600 python tools/prof_run.py caustics_bdpt 1 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['workload'], round(d['ms'],1), {k:round(v[0],1) for k,v in d['kernel_ms'].items()}, d['counters']['closest'])"
done
