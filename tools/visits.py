"""Visit counters per ray (closest-hit and occlusion kernels) for a workload: the N_* of DESIGN.md's byte formula."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from lumo_b200 import native
name = sys.argv[1]; spp = int(sys.argv[2]) if len(sys.argv) > 2 else 1
prog, blob, integrator, _ = bench.build_workload(name)
ctx = native.GpuContext(0); G = native.GpuScene(ctx, blob)
ctx.count_visits(True)
px, sp, cnt, _, ms = G.render(integrator=integrator, spp=spp, seed=1, rr_delta=0.05)
vc, vo = ctx.visits()
ctx.count_visits(False)
print(json.dumps({"workload": name, "closest_rays": cnt["closest"], "occlusion_rays": cnt["occlusion"],
                  "closest_per_ray": {k: round(v / max(cnt["closest"], 1), 2) for k, v in vc.items()},
                  "occlusion_per_ray": {k: round(v / max(cnt["occlusion"], 1), 2) for k, v in vo.items()}}))
