#!/bin/bash
mkdir -p gpurun_out
PROF_WARM=0 python tools/prof_run.py bunny 4 > gpurun_out/prof5_plain.log 2>&1 &&
PROF_WARM=0 ncu --set full --clock-control none --import-source on -k regex:"k_wave_trace|k_wave_occlude" -s 2 -c 6 -o gpurun_out/prof5 python tools/prof_run.py bunny 4 > gpurun_out/ncu_full5.log 2>&1
tail -1 gpurun_out/ncu_full5.log
python - <<'PY'
import time, sys, numpy as np
sys.path.insert(0, ".")
import bench
from lumo_b200 import native
prog, blob, ig, spp = bench.build_workload("bunny")
ctx = native.GpuContext(0)
for i in range(3):
    t0 = time.perf_counter(); sc = native.GpuScene(ctx, blob); t1 = time.perf_counter()
    px, sp, cnt, _, ms = sc.render(integrator=ig, spp=spp, seed=1 + i); t2 = time.perf_counter()
    sc.close(); t3 = time.perf_counter()
    print("e2e pieces: upload %.1f ms, render call %.1f ms (device %.1f ms), close %.1f ms" % (1e3 * (t1 - t0), 1e3 * (t2 - t1), ms, 1e3 * (t3 - t2)))
PY
