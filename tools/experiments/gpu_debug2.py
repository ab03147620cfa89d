import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import conftest, oracle_lib
from lumo_b200 import native, scenes, PixelFilter
ctx = native.GpuContext(0)
s, cam, ig = scenes.cornell(resolution=(64, 64))
cam._pixel_filter = PixelFilter.square(0.5)
prog = s._program(cam); blob = native.build_blob(prog)
O = oracle_lib.OracleScene(prog); G = native.GpuScene(ctx, blob)
epx, _, ecnt, _ = O.render(integrator=0, spp=1, seed=7, rng_mode=1, rr_delta=0.02, sampler=0, threads=1)
gpx, _, gcnt, _, ms = G.render(integrator=0, spp=1, seed=7, rr_delta=0.02, sampler=0)
d = np.abs(gpx[..., :3] - epx[..., :3]).max(-1); tol = 1e-9 * (np.abs(epx[..., :3]).max(-1) + 1e-6)
bad = np.argwhere(d > tol)
print("differing paths", len(bad), "of", d.size)
for y, x in bad[:3]:
    pix = int(x + y * 64)
    print("=== pixel", x, y, pix, epx[y, x], gpx[y, x]); sys.stdout.flush()
    os.environ["ORACLE_DEBUG_PIXEL"] = str(pix); os.environ["LUMO_DEBUG_PIXEL"] = str(pix)
    O.render(integrator=0, spp=1, seed=7, rng_mode=1, rr_delta=0.02, sampler=0, threads=1)
    sys.stdout.flush()
    G.render(integrator=0, spp=1, seed=7, rr_delta=0.02, sampler=0)
    sys.stdout.flush()
