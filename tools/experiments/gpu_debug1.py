import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import conftest, oracle_lib
from lumo_b200 import native
ctx = native.GpuContext(0)
for name, ig in [("cornell", 0), ("cornell", 1), ("bunny", 0)]:
    prog, blob, _ = conftest.small_scene(name)
    O = oracle_lib.OracleScene(prog); G = native.GpuScene(ctx, blob)
    for sampler in (0, 2):
        epx, _, ecnt, _ = O.render(integrator=ig, spp=1, seed=7, rng_mode=1, rr_delta=0.02, sampler=sampler)
        gpx, _, gcnt, _, ms = G.render(integrator=ig, spp=1, seed=7, rr_delta=0.02, sampler=sampler)
        print(name, ig, "sampler", sampler, ecnt, gcnt, "ms", ms)
        dw = np.abs(gpx[..., 3] - epx[..., 3]); print("  weight max diff", dw.max())
        d = np.abs(gpx[..., :3] - epx[..., :3]).max(-1); tol = 1e-9 * (np.abs(epx[..., :3]).max(-1) + 1e-6)
        bad = np.argwhere(d > tol)
        print("  differing pixels", len(bad), "of", d.size)
        for y, x in bad[:12]:
            print("   ", x, y, epx[y, x], gpx[y, x])
