"""A/B timing of liblumo_gpu.so variants (tools/build_variant.sh) on the GPU box: python tools/ab_run.py <workload> <spp> [variant ...]
Each variant renders in its own process (LUMO_GPU_SO), three times; prints the per-class kernel times of the best run."""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, ROOT)
    import bench
    from lumo_b200 import native
    name, spp = sys.argv[2], int(sys.argv[3])
    prog, blob, integrator, _ = bench.build_workload(name)
    ctx = native.GpuContext(0); sc = native.GpuScene(ctx, blob)
    best = None
    for k in range(3):
        _, _, cnt, _, ms = sc.render(integrator=integrator, spp=spp, seed=3, rr_delta=0.05)
        kt = ctx.kernel_times()
        if best is None or ms < best[0]: best = (ms, {a: round(b[0], 1) for a, b in kt.items()}, cnt["closest"] + cnt["occlusion"])
    print(json.dumps({"ms": round(best[0], 1), "kernel_ms": best[1], "mrays_per_s": round(best[2] / best[0] / 1e3, 1)}))
    sys.exit(0)
name, spp = sys.argv[1], sys.argv[2]
variants = sys.argv[3:] or ["default"]
for v in variants:
    env = dict(os.environ)
    if v != "default": env["LUMO_GPU_SO"] = os.path.join(ROOT, "lumo_b200", "variants", "liblumo_gpu_%s.so" % v)
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", name, spp], env=env, capture_output=True, text=True)
    print("%-24s %s" % (v, r.stdout.strip().splitlines()[-1] if r.returncode == 0 and r.stdout.strip() else "FAILED: " + r.stderr[-300:]), flush=True)
