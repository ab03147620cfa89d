"""Small renders through every kernel family (PT, DL, BDPT, textures, batch API, small BDPT batches, node-major traversal): a quick does-everything-run check."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from conftest import small_scene
from lumo_b200 import native
ctx = native.GpuContext(0)
for name in ("cornell", "textured", "caustics", "bistro"):
    prog, blob, ig = small_scene(name)
    G = native.GpuScene(ctx, blob)
    for integ in (0, 1, 2):
        if name == "bistro" and integ == 2: continue
        r = G.render(integrator=integ, spp=1, seed=3)
        print(name, integ, r[2]["camera_paths"], r[2]["closest"], r[2]["occlusion"], float(np.nan_to_num(r[0]).sum()))
    rs = np.random.RandomState(1)
    o = rs.rand(2000, 3) * 2 - 1; d = rs.randn(2000, 3); d /= np.linalg.norm(d, axis=1, keepdims=True)
    G.trace_closest(o, d); G.trace_any(o, d, np.full(2000, 5.0)); G.trace_first_found(o, d)
    G.close()
os.environ["LUMO_BDPT_BATCH"] = "1024"
prog, blob, ig = small_scene("caustics")
G = native.GpuScene(ctx, blob); print("small batches", G.render(integrator=2, spp=2, seed=4)[2]["iterations"]); G.close()
del os.environ["LUMO_BDPT_BATCH"]
os.environ["LUMO_TRACE_NM"] = "1"
c2 = native.GpuContext(0)
del os.environ["LUMO_TRACE_NM"]
prog, blob, ig = small_scene("bunny")
G = native.GpuScene(c2, blob); print("nm", G.render(integrator=0, spp=64, seed=4)[2]["closest"]); G.close(); c2.close()
ctx.close()
print("done")
