#!/bin/bash
python - <<'PY'
import sys; sys.path.insert(0, "."); sys.path.insert(0, "tests")
import bench, numpy as np
from lumo_b200 import native
prog, blob, ig, spp = bench.build_workload("caustics_bdpt")
ctx = native.GpuContext(0); G = native.GpuScene(ctx, blob)
G.render(integrator=0, spp=1, seed=2, rr_delta=0.05)
px, sp, cnt, _, ms = G.render(integrator=0, spp=2, seed=1, rr_delta=0.05)
kt = ctx.kernel_times()
print("caustics PT 2spp", round(ms, 1), {k: round(v[0], 1) for k, v in kt.items()}, cnt["closest"], cnt["occlusion"], "trace Mrays/s", round(cnt["closest"] / kt["trace"][0] / 1e3, 1), "iters", cnt["iterations"])
print(ctx.iter_log()[:40])
PY
LUMO_BDPT_LOG=1 python tools/prof_run.py caustics_bdpt 1 2>&1 | grep "^\[bdpt\]" | tail -10
