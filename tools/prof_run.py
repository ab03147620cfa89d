"""Profiling driver: one short render of a workload through the C ABI (no oracle, no CPU leg)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lumo_b200 import native
import bench
name = sys.argv[1] if len(sys.argv) > 1 else "bunny"
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 1
wave = int(sys.argv[3]) if len(sys.argv) > 3 else 0
prog, blob, integrator, _ = bench.build_workload(name)
ctx = native.GpuContext(0)
G = native.GpuScene(ctx, blob)
if os.environ.get("PROF_WARM", "1") == "1":
    G.render(integrator=integrator, spp=1, seed=2, rr_delta=0.05, wave_paths=wave)      # module load + first-launch costs
px, sp, cnt, _, ms = G.render(integrator=integrator, spp=spp, seed=1, rr_delta=0.05, wave_paths=wave)
print(json.dumps({"workload": name, "spp": spp, "ms": ms, "counters": cnt, "kernel_ms": ctx.kernel_times(), "iter_log": ctx.iter_log()[:64]}))
G.close(); ctx.close()
