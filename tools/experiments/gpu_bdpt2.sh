#!/bin/bash
for c in 131072 262144 524288 1048576; do
  echo -n "batch $c: "; LUMO_BDPT_BATCH=$c timeout 600 python tools/prof_run.py caustics_bdpt 1 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['workload'], round(d['ms'],1), {k:round(v[0],1) for k,v in d['kernel_ms'].items()})"
done
