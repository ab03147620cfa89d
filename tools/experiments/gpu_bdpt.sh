#!/bin/bash
for s in "caustics_bdpt 1" "cornell 4"; do
  timeout 600 python tools/prof_run.py $s 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['workload'], round(d['ms'],1), {k:round(v[0],1) for k,v in d['kernel_ms'].items()}, 'rays', d['counters']['closest']+d['counters']['occlusion'], d['counters'])"
done
python - <<'PY'
import sys; sys.path.insert(0, "."); sys.path.insert(0, "tests")
import bench, numpy as np
from lumo_b200 import native
prog, blob, ig, spp = bench.build_workload("bunny")
ctx = native.GpuContext(0); G = native.GpuScene(ctx, blob)
G.render(integrator=2, spp=1, seed=2)
px, sp, cnt, _, ms = G.render(integrator=2, spp=1, seed=1)
print("bunny bdpt 1spp", round(ms, 1), ctx.kernel_times(), cnt)
PY
bash tools/gpu_tests.sh -k "bdpt or relmse or ranges"
