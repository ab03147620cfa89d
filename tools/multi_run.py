"""lumo_gpu_render_multi from one process on every GPU of the box: python tools/multi_run.py [workload] [spp per GPU]
Prints one JSON line per GPU count (1, 2, ... all): wall-clock Mrays/s of the call (host film buffers included)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
from lumo_b200 import native
import bench

name = sys.argv[1] if len(sys.argv) > 1 else "bunny"
spp_per_gpu = int(sys.argv[2]) if len(sys.argv) > 2 else 8
prog, blob, integrator, _ = bench.build_workload(name)
nd = C.c_int32(0); native._check(native.gpu_lib().lumo_gpu_device_count(C.byref(nd)), "device_count")
ns = [n for n in (1, 2, 4, 8) if n <= nd.value]
ctxs = [native.GpuContext(g) for g in range(nd.value)]
scenes = [native.GpuScene(c, blob) for c in ctxs]
ref = None
for n in ns:
    best = None
    for rep in range(3):
        t0 = time.perf_counter()
        px, sp, cnt, _, ms = native.render_multi(scenes[:n], integrator=integrator, spp=spp_per_gpu * n, seed=1)
        dt = time.perf_counter() - t0
        if rep and (best is None or dt < best[0]): best = (dt, cnt, ms)
    dt, cnt, ms = best
    print(json.dumps({"workload": name, "n_gpus": n, "spp": spp_per_gpu * n, "wall_ms": 1e3 * dt, "device_ms": ms,
                      "mrays_per_s_wall": (cnt["closest"] + cnt["occlusion"]) / dt / 1e6, "camera_paths": cnt["camera_paths"], "film_mean": float(px[..., :3].sum() / px[..., 3].sum())}), flush=True)
for s in scenes: s.close()
for c in ctxs: c.close()
