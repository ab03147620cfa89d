#!/bin/bash
# short renders of several workloads: total and per-stage device ms
for s in "bunny 4" "bistro 1" "conference 4" "cornell 8" "textured 4" "caustics_bdpt 1"; do
  timeout 300 python tools/prof_run.py $s 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['workload'], round(d['ms'],1), {k:round(v[0],1) for k,v in d['kernel_ms'].items()}, 'rays', d['counters']['closest']+d['counters']['occlusion'])"
done
