"""Same-stream film parity, GPU vs oracle (counter mode), for every small config and integrator: how many pixels / counters
differ at all.  Run on the GPU box: python tools/parity_report.py > gpurun_out/parity.txt"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib
from conftest import small_scene, SMALL
from lumo_b200 import native

ctx = native.GpuContext(0)
def rgb(px):
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.nan_to_num(px[..., :3] / px[..., 3:4])
for name in SMALL:
    for integrator, spp in ((0, 4), (1, 4), (2, 2)):
        prog, blob, _ = small_scene(name, box_filter=True)
        O = oracle_lib.OracleScene(prog); G = native.GpuScene(ctx, blob)
        e = O.render(integrator=integrator, spp=spp, seed=7, rng_mode=1)
        g = G.render(integrator=integrator, spp=spp, seed=7)
        er, gr = rgb(e[0]), rgb(g[0])
        d = np.abs(gr - er)
        bad9 = (d > 1e-9 * (np.abs(er) + 1e-6)).any(axis=-1).mean()
        bad12 = (d > 1e-12 * (np.abs(er) + 1e-6)).any(axis=-1).mean()
        sbad = (np.abs(g[1] - e[1]) > 1e-9 * (np.abs(e[1]) + 1e-6)).any(axis=-1).mean()
        cnt = {k: (g[2][k], e[2][k]) for k in ("closest", "occlusion", "cost")}
        dl = np.abs(g[3] / e[3] - 1).max() if integrator != 1 else 0.0
        print("%-10s integrator %d spp %d: pixels differing >1e-9: %.5f  >1e-12: %.5f  splat pixels >1e-9: %.5f  delta max rel %.2e  counters (gpu, oracle) %s" % (name, integrator, spp, bad9, bad12, sbad, dl, cnt), flush=True)
        G.close(); O.close()
ctx.close()
