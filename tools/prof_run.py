"""One short render for profiling: python tools/prof_run.py <workload> <spp> [integrator]  (ncu wraps this)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from lumo_b200 import native
name = sys.argv[1]; spp = int(sys.argv[2])
prog, blob, integrator, _ = bench.build_workload(name)
if len(sys.argv) > 3: integrator = int(sys.argv[3])
ctx = native.GpuContext(0)
sc = native.GpuScene(ctx, blob)
_, _, cnt, _, ms = sc.render(integrator=integrator, spp=spp, seed=3, rr_delta=float(os.environ.get("PROF_RR_DELTA", "0")))
print(name, spp, "spp:", ms, "ms", cnt, ctx.kernel_times())
if os.environ.get("PROF_ITERLOG"):
    import json
    json.dump({"workload": name, "spp": spp, "iterations": ctx.iter_log(), "counters": cnt}, open(os.environ["PROF_ITERLOG"], "w"))
sc.close(); ctx.close()
