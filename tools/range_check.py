"""G6 on one GPU: the films of N disjoint sample ranges summed vs one render of the whole range (what bench.py's film_check
does across ranks).  python tools/range_check.py <workload> <ranks> <spp per rank> [seed]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from lumo_b200 import native
name, n, k = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]); seed = int(sys.argv[4]) if len(sys.argv) > 4 else 77
prog, blob, integrator, _ = bench.build_workload(name)
ctx = native.GpuContext(0); sc = native.GpuScene(ctx, blob)
parts = None; cnts = []
for r in range(n):
    px, sp, cnt, dl, ms = sc.render(integrator=integrator, seed=seed, spp_begin=r * k, spp_end=(r + 1) * k, total_spp=n * k)
    parts = px.copy() if parts is None else parts + px; cnts.append(cnt)
full, _, cf, dlf, _ = sc.render(integrator=integrator, seed=seed, spp_begin=0, spp_end=n * k, total_spp=n * k)
d = np.abs(parts - full); den = np.abs(full).max()
print("max abs diff", d.max(), "max value", den, "rel", d.max() / den)
print("closest", sum(c["closest"] for c in cnts), cf["closest"], "occlusion", sum(c["occlusion"] for c in cnts), cf["occlusion"], "cost", sum(c["cost"] for c in cnts), cf["cost"])
bad = np.argwhere(d > 1e-9 * den)
print("pixels differing:", len(bad))
for y, x, c in bad[:10]: print((int(x), int(y), int(c)), parts[y, x, c], full[y, x, c])
