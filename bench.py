#!/usr/bin/env python
"""bench.py — Mrays/s and Msamples/s of the lumo hot path on B200 (BASELINE.json metric).

Workload at N = 1: `bistro` — examples/bistro.rs, PathTrace, 1920x1080, the configuration BASELINE.json's target is
quoted on (synthetic 1.05 M-triangle street in 1024 kd-trees, 4097 lights: the assets do not exist offline).  A "step"
is one pass of the hot path over one batch: `--spp` samples per pixel (default 32) of the whole image through the
wavefront pipeline (camera rays -> closest-hit traversal -> shading / NEE -> occlusion traversal -> film).
rays = closest-hit queries + occlusion queries, counted on the device.  With N GPUs (torchrun) every rank renders its
own sample range of an N*spp render with the scene replicated (weak scaling) and the film accumulators are combined with
ONE NCCL reduce inside the timed region.

  value         whole-job Mrays/s, scene resident in HBM, film left on the device (CUDA events, max over ranks)
  e2e           the same metric through the public calls with HOST buffers, every step: scene blob upload from pinned host
                memory (H2D) + render (+ NCCL reduce at N > 1) + film download into pinned host memory (D2H, 56 B/pixel)
  roofline      the kernel class that takes the most time of the step; `roofline_by_kernel` has all three (closest-hit
                trace, shading, occlusion) — algorithmic bytes (DESIGN.md byte formulas x device visit counters from an
                untimed counting pass of the same step) / CUDA-event time of the class
  cpu_baseline  the C++ restatement of lumo's CPU renderer (oracle/, reference schedule, all host cores) at the SAME spp on
                every k-th 16x16 tile of the same image (k chosen for ~10-20 s of CPU work)
  other_workloads  the four other BASELINE configs, each measured the same way with fewer steps
`--impl reference` runs only that CPU restatement (the reference is Rust and cannot be built here), same config."""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {   # name -> (scene kwargs, integrator, default spp per step)
    "bistro": (dict(), 0, 32),
    "bunny": (dict(), 0, 32),
    "cornell": (dict(), 0, 64),
    "dragon": (dict(), 0, 16),
    "caustics_bdpt": (dict(), 2, 4),
    "conference": (dict(), 0, 32),
    "conference_dl": (dict(), 1, 32),
    "textured": (dict(n_tris=20000, resolution=(1024, 768)), 0, 16),
}
DESCR = {
    "bunny": "examples/bunny.rs PathTrace 1024x768, synthetic 69k-triangle stand-in mesh (assets are not available offline)",
    "cornell": "examples/cornell.rs PathTrace 512x512",
    "dragon": "examples/dragon.rs PathTrace 1024x768, synthetic 870k-triangle stand-in mesh",
    "caustics_bdpt": "examples/caustics.rs BDPathTrace 1024x768, synthetic 15k-triangle stand-in mesh instanced twice (mirror + dispersive glass)",
    "conference": "examples/conference.rs PathTrace 1024x768, synthetic 332k-triangle room in 64 kd-trees",
    "conference_dl": "examples/conference.rs DirectLight 1024x768, synthetic 332k-triangle room in 64 kd-trees",
    "bistro": "examples/bistro.rs PathTrace 1920x1080, synthetic 1.05M-triangle street in 1024 kd-trees, 4097 lights",
    "textured": "not a reference example: empty box with every Texture kind, bump maps, image light + environment map (scenes.textured) PathTrace 1024x768",
}
INTEGRATORS = ["PathTrace", "DirectLight", "BDPathTrace"]
TARGET_MRAYS = 1000.0   # BASELINE.json north_star: >= 1 Gray/s per B200 on Bistro path tracing


def build_workload(name):
    from lumo_b200 import scenes, native
    kw, integrator, spp = WORKLOADS[name]
    s, cam, _ = scenes.CONFIGS[name.split("_")[0]](**kw)
    prog = s._program(cam)
    blob = native.build_blob(prog)
    return prog, blob, integrator, spp


def config_of(name, integrator, spp, world, res):
    """The `config` object — identical in the repo arm and the reference arm."""
    return {"workload": name, "description": DESCR[name], "integrator": INTEGRATORS[integrator], "resolution": list(res), "spp_per_step_per_gpu": spp,
            "parallelism": "scene replicated, samples sharded, 1 NCCL reduce" if world > 1 else "single GPU",
            "l2": "256 MiB buffer written between timed steps (L2 flush); wave state and film also exceed the 126 MB L2"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.lines, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100", "-i", str(index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try: self.proc.wait(timeout=2)
        except Exception: self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 7: continue
            try: sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError: continue
            for k, n in enumerate(names):
                if f[3 + k].lower().startswith("active"): reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def ray_bytes(v, n_rays, fixed):
    """DESIGN.md byte formula of the reference-order traversal: per-ray queue traffic + nodes / instances / leaf entries / triangles visited."""
    return fixed * n_rays + 64 * v["tlas_nodes"] + 96 * v["inst"] + 16 * v["kd_nodes"] + 4 * v["leaf_idx"] + 72 * v["tri_tests"] + 16 * v["sphere_tests"]


def occl_bytes(st, vis, n_rays):
    """Occlusion pipeline: 60 B shadow-queue entry per ray, 128 B per BVH node, 8 B per leaf primitive, 72 B per triangle test,
    16 B per sphere test, 8 B per confirm-queue entry, plus the confirming traversal's own visits (reference-order formula)."""
    return 60 * n_rays + 128 * st["nodes"] + 8 * st["prims"] + 72 * st["tri_tests"] + 16 * st["sphere_tests"] + 8 * st["candidates"] + ray_bytes(vis, 0, 0) + 60 * st["candidates"]


def micro_trace(scene, dev, torch, np, n=1 << 22):
    """lumo_gpu_trace_closest_dev on device-resident batches; kernel time from the library's CUDA events."""
    P = scene.blob.params; cam = P["camera"]
    W, H = int(cam["res_x"]), int(cam["res_y"])
    g = torch.Generator(device=dev); g.manual_seed(23)
    rs = np.random.RandomState(23)
    raster = np.stack([rs.rand(n) * W, rs.rand(n) * H, np.zeros(n), np.ones(n)], 0)
    def xf(m16, v):
        r = np.asarray(m16, np.float64).reshape(4, 4) @ v
        w = np.where(r[3] == 0.0, 1.0, r[3]); r = r / w; r[3] = 1.0
        return r
    c = xf(cam["camera_to_screen_inv"], xf(cam["screen_to_raster_inv"], raster))
    d_local = c[:3] / np.linalg.norm(c[:3], axis=0)
    m = np.asarray(cam["world_to_camera_inv"], np.float64).reshape(4, 4)
    o_w = np.repeat(m[:3, 3:4], n, axis=1)
    d_w = m[:3, :3] @ d_local; d_w = d_w / np.linalg.norm(d_w, axis=0)
    out = {}
    lo = torch.tensor(np.array(P["bounds_lo"]), dtype=torch.float64, device=dev); hi = torch.tensor(np.array(P["bounds_hi"]), dtype=torch.float64, device=dev)
    o2 = lo + torch.rand((n, 3), dtype=torch.float64, device=dev, generator=g) * (hi - lo)
    z = 1 - 2 * torch.rand(n, dtype=torch.float64, device=dev, generator=g); ph = 2 * np.pi * torch.rand(n, dtype=torch.float64, device=dev, generator=g)
    r = torch.sqrt(torch.clamp(1 - z * z, min=0))
    d2 = torch.stack([r * torch.cos(ph), r * torch.sin(ph), z], -1).contiguous()
    batches = {"primary": (torch.tensor(np.ascontiguousarray(o_w.T), device=dev), torch.tensor(np.ascontiguousarray(d_w.T), device=dev)), "incoherent": (o2.contiguous(), d2)}
    obj = torch.empty(n, dtype=torch.int32, device=dev); tri = torch.empty(n, dtype=torch.int32, device=dev)
    t = torch.empty(n, dtype=torch.float64, device=dev); bary = torch.empty((n, 2), dtype=torch.float64, device=dev)
    for name, (o, d) in batches.items():
        best = None
        for k in range(4):
            ms = scene.trace_closest_dev(o.data_ptr(), d.data_ptr(), n, obj.data_ptr(), tri.data_ptr(), t.data_ptr(), bary.data_ptr())
            if k > 0: best = ms if best is None else min(best, ms)
        out[name] = {"mrays_per_s": n / best / 1e3, "ms": best, "rays": n, "hit_fraction": float((obj != -1).double().mean().item())}
    return out


# ---- CPU arm: the C++ restatement of lumo's CPU renderer ------------------------------------------------------------------
class CpuArm:
    """lumo's CPU path (C++ restatement, reference schedule: 16x16 tiles x 256-sample batches from a shared queue, per-tile
    xorshift and ring-buffer RR threshold; renderer.rs:174-204, task.rs:25-81), -O3 -march=native, all host cores, at the
    workload's own spp on every `tile_step`-th tile of the full-resolution image."""
    def __init__(self, prog, integrator, spp, target_seconds):
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib
        self.O = oracle_lib.OracleScene(prog, native=True)
        self.integrator, self.spp = integrator, spp
        self.threads = os.cpu_count() or 1
        n_tiles = ((self.O.res_x + 15) // 16) * ((self.O.res_y + 15) // 16)
        # calibrate on ~2 tiles per thread, then choose the stride for about target_seconds per step
        cal = max(1, n_tiles // (2 * self.threads))
        t0 = time.time(); _, _, cnt, _ = self.O.render(integrator=integrator, spp=spp, seed=99, rng_mode=0, threads=self.threads, tile_step=cal); dt = max(time.time() - t0, 1e-3)
        tiles_cal = -(-n_tiles // cal)
        per_tile = dt / tiles_cal * min(self.threads, tiles_cal) / self.threads       # wall seconds per tile at full occupancy of the cores
        want_tiles = max(self.threads * 2, int(target_seconds / max(per_tile, 1e-9)))
        self.tile_step = max(1, n_tiles // min(n_tiles, want_tiles))
        self.n_tiles, self.tiles = n_tiles, -(-n_tiles // self.tile_step)

    def step(self, seed):
        t0 = time.time()
        _, _, cnt, _ = self.O.render(integrator=self.integrator, spp=self.spp, seed=seed, rng_mode=0, threads=self.threads, tile_step=self.tile_step)
        dt = time.time() - t0
        return cnt, dt

    def sample_text(self):
        return ("%d spp (the workload's own) on every %d-th 16x16 tile of the full-resolution image: %d of %d tiles per step (lumo CPU path, C++ restatement with the "
                "reference tile/batch schedule and RR recurrence; the Rust reference cannot be built offline)" % (self.spp, self.tile_step, self.tiles, self.n_tiles))

    def close(self):
        self.O.close()


def cpu_measure(prog, integrator, spp, target_seconds, steps=1, warmup=0, seed0=1):
    arm = CpuArm(prog, integrator, spp, target_seconds)
    rays = paths = secs = 0.0
    for i in range(warmup + steps):
        cnt, dt = arm.step(seed0 + i)
        if i >= warmup:
            rays += cnt["closest"] + cnt["occlusion"]; paths += cnt["camera_paths"]; secs += dt
    out = {"value": rays / secs / 1e6, "unit": "Mrays/s", "cores": arm.threads, "kind": "port", "msamples_per_s": paths / secs / 1e6,
           "rays_per_sample": rays / max(paths, 1), "seconds_per_step": secs / steps, "sample": arm.sample_text()}
    arm.close()
    return out


_RESULT_OUT = None


def _guard_stdout():
    """stdout carries exactly ONE JSON line.  Libraries write there too (NCCL prints its version banner on communicator
    creation, `make` may echo): keep the real stdout for the result and point fd 1 at stderr for everybody else."""
    global _RESULT_OUT
    if _RESULT_OUT is None:
        sys.stdout.flush()
        _RESULT_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
    return _RESULT_OUT


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    prog, blob, integrator, spp = build_workload(args.workload)
    if args.spp: spp = args.spp
    from lumo_b200 import native
    B = native.Blob(blob); res = (int(B.params["camera"]["res_x"]), int(B.params["camera"]["res_y"]))
    # each step a bounded sample: sized so that warmup + steps end within a few minutes
    per_step = max(2.0, min(15.0, 240.0 / max(args.steps + args.warmup, 1)))
    r = cpu_measure(prog, integrator, spp, per_step, steps=args.steps, warmup=args.warmup)
    line = {"impl": "reference", "metric": "Mrays/s", "value": r["value"], "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * r["seconds_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(args.workload, integrator, spp, int(os.environ.get("WORLD_SIZE", "1")), res),
            "msamples_per_s": r["msamples_per_s"], "rays_per_sample": r["rays_per_sample"],
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), file=_guard_stdout(), flush=True)
    return 0


# ---- GPU arm ----------------------------------------------------------------------------------------------------------------
def measure(name, args, env, steps, warmup, full):
    """One workload on this rank's GPU (and, at world > 1, together with the other ranks).  Returns the JSON line's content on
    rank 0, None elsewhere.  full: also micro batches / film kernel / strong-scaling extras (main workload only)."""
    torch, np, native, dist = env["torch"], env["np"], env["native"], env["dist"]
    world, rank, dev, ctx, stream, flush = env["world"], env["rank"], env["dev"], env["ctx"], env["stream"], env["flush"]
    prog, blob, integrator, spp = build_workload(name)
    if args.spp and name == args.workload: spp = args.spp
    strong = bool(args.total_spp) and name == args.workload
    if strong:
        assert args.total_spp % world == 0, "--total-spp must be a multiple of the number of GPUs"
        spp = args.total_spp // world
    total_spp = spp * world
    s0, s1 = rank * spp, (rank + 1) * spp
    ctx.set_stream(stream.cuda_stream)
    scene = native.GpuScene(ctx, blob)
    W, H = scene.res_x, scene.res_y
    film = torch.zeros(W * H * 7, dtype=torch.float64, device=dev)      # pixels[W*H*4] then splats[W*H*3]
    px_ptr, sp_ptr = film.data_ptr(), film.data_ptr() + W * H * 4 * 8

    def step(seed, a=s0, b=s1, total=total_spp, reduce=True):
        cnt, ms = scene.render_dev(px_ptr, sp_ptr, integrator=integrator, seed=seed, spp_begin=a, spp_end=b, total_spp=total, wave_paths=args.wave_paths)
        if dist is not None and reduce:
            dist.reduce(film, dst=0)       # the one collective of the path: film accumulators over NVLink
        return cnt

    def barrier():
        if dist is not None: dist.barrier()
        torch.cuda.synchronize()

    for i in range(warmup):
        step(100 + i)
    barrier()
    clocks = ClockSampler(env["local"]) if rank == 0 else None
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    tot = {"closest": 0, "occlusion": 0, "camera_paths": 0, "cost": 0, "gpu_launches": 0, "iterations": 0}
    ktimes = {"regen": [0.0, 0], "trace": [0.0, 0], "shade": [0.0, 0], "occlude": [0.0, 0]}
    seeds = [1 + i for i in range(steps)]
    barrier()
    for i in range(steps):
        flush.zero_()                      # L2 flush between timed iterations (outside the events)
        ev[i][0].record(stream)
        cnt = step(seeds[i])
        ev[i][1].record(stream)
        for k in tot: tot[k] += cnt[k]
        for k, (ms, n) in ctx.kernel_times().items(): ktimes[k][0] += ms; ktimes[k][1] += n
    barrier()
    clk = clocks.stop() if clocks else None
    ms_total = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    c = torch.tensor([tot["closest"], tot["occlusion"], tot["camera_paths"], tot["gpu_launches"]], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX); dist.all_reduce(c, op=dist.ReduceOp.SUM)
    ms_total = float(t.item()); closest, occl, paths, launches = (float(x) for x in c.tolist())
    rays = closest + occl
    value = rays / (ms_total * 1e-3) / 1e6

    # ---- e2e: host buffers on both sides of every step, the same definition at every N ---------------------------------------
    # upload of the scene blob from pinned host memory, render, NCCL reduce (N > 1), download of the 56 B/pixel film accumulators
    # into pinned host memory on rank 0; wall clock, max over ranks, one discarded step then up to 5 timed ones
    pin_blob = torch.frombuffer(bytearray(blob), dtype=torch.uint8).pin_memory()
    pin_film = torch.empty(W * H * 7, dtype=torch.float64).pin_memory() if rank == 0 else None
    n_e2e = max(1, min(steps, 5 if full else 2))
    e_rays, e_t = 0.0, 0.0
    for i in range(1 + n_e2e):
        barrier(); t0 = time.perf_counter()
        sc2 = native.GpuScene(ctx, blob, host_ptr=pin_blob.data_ptr())
        cnt2, _ = sc2.render_dev(px_ptr, sp_ptr, integrator=integrator, seed=seeds[min(i, len(seeds) - 1)], spp_begin=s0, spp_end=s1, total_spp=total_spp, wave_paths=args.wave_paths)
        if dist is not None: dist.reduce(film, dst=0)
        if rank == 0: pin_film.copy_(film, non_blocking=True)
        torch.cuda.synchronize()
        sc2.close()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        r2 = torch.tensor([cnt2["closest"] + cnt2["occlusion"]], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX); dist.all_reduce(r2, op=dist.ReduceOp.SUM)
        if i > 0: e_rays += float(r2.item()); e_t += float(dt.item())
    e2e = {"value": e_rays / e_t / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": world * (len(blob) + 64), "d2h_bytes_per_step": W * H * 56 + 64 * world, "steps": n_e2e,
           "what": "per step and per rank: lumo_gpu_scene_upload from a pinned host blob + lumo_gpu_render_dev" + (" + ONE NCCL reduce of the film" if world > 1 else "") +
                   " + download of the 56 B/pixel film accumulators into pinned host memory on rank 0 + lumo_gpu_scene_destroy; wall clock, max over ranks, after one discarded step"}

    # ---- G6 inside the run (N > 1): the reduced film of N ranks equals rank 0's own render of the whole range ---------------
    film_check = None
    if dist is not None and full:
        k = 2                                                     # 2 spp per rank keeps the single-GPU side short
        scene.render_dev(px_ptr, sp_ptr, integrator=integrator, seed=77, spp_begin=rank * k, spp_end=(rank + 1) * k, total_spp=k * world)
        dist.reduce(film, dst=0)
        if rank == 0:
            multi = film.clone()
            scene.render_dev(px_ptr, sp_ptr, integrator=integrator, seed=77, spp_begin=0, spp_end=k * world, total_spp=k * world)
            den = float(multi.abs().max().item())
            diff = float((multi - film).abs().max().item())
            film_check = {"spp_per_rank": k, "max_abs_diff": diff, "max_abs_value": den, "max_rel_diff": diff / max(den, 1e-300), "bound": 1e-12,
                          "ok": diff <= 1e-12 * max(den, 1e-300), "checksum": float(multi.sum().item()),
                          "what": "film accumulators of the N-rank render after the NCCL reduce vs rank 0's own render of the same %d samples per pixel (same seed)" % (k * world)}
            del multi
        dist.barrier()
    if rank != 0:
        scene.close(); del film
        return None

    # ---- rank 0 only from here: counting pass, rooflines, cpu baseline ------------------------------------------------------------
    line = {"metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_total / steps, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(config_of(name, integrator, spp, world, (W, H)), **({"total_spp_per_step": total_spp} if strong else {})),
            "msamples_per_s": paths / (ms_total * 1e-3) / 1e6, "rays_per_sample": rays / max(paths, 1),
            "rays": {"closest_hit": closest, "occlusion": occl, "camera_paths": paths, "reference_style_cost": tot["cost"]},
            "gpu_launches": int(launches), "clocks": clk, "e2e": e2e}
    if name == "bistro":
        line["target"] = {"what": "BASELINE.json north_star: >= 1 Gray/s per B200 on Bistro path tracing", "mrays_per_s_per_gpu": TARGET_MRAYS, "achieved_fraction": value / world / TARGET_MRAYS}
    if film_check is not None: line["film_check"] = film_check
    kms = {k: v[0] / steps for k, v in ktimes.items()}
    step_ms = ms_total / steps
    line["kernel_ms_per_step"] = dict(kms, outside_kernel_classes=max(step_ms - sum(kms.values()), 0.0), covered_fraction=sum(kms.values()) / step_ms,
                                      note="device time between CUDA events around each kernel class of the main pass on this rank; `outside` = the two pilot rounds of the RR threshold, memsets, queue bookkeeping, host synchronisation every 4 wave iterations")

    # one untimed counting pass of the last timed step (visit-counting instantiations of the traversal kernels)
    ctx.count_visits(True)
    cntc, _ = scene.render_dev(px_ptr, sp_ptr, integrator=integrator, seed=seeds[-1], spp_begin=s0, spp_end=s1, total_spp=total_spp, wave_paths=args.wave_paths)
    vis_closest, vis_occl = ctx.visits()
    ost = ctx.occlusion_stats(); cst = ctx.closest_stats()
    ctx.count_visits(False)
    peaks = {}
    try: peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception: pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)"
    traffic_db = {}
    try: traffic_db = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(name, {})
    except Exception: pass
    bounces = max(cntc["closest"], 1)
    n_shadow_per_bounce = cntc["occlusion"] / bounces
    by = {}

    def entry(kernel, cls, bytes_total, units, unit_name, note):
        ms, n = kms[cls], ktimes[cls][1] / steps
        if ms <= 0: return None
        ach = bytes_total / (ms * 1e-3) / 1e9
        e = {"bound": "hbm", "kernel": kernel, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "peak_source": peak_src,
             "algorithmic_bytes_per_launch": bytes_total / max(n, 1), "launches_per_step": n, "avg_launch_ms": ms / max(n, 1), "ms_per_step": ms,
             "bytes_per_" + unit_name: bytes_total / max(units, 1), "traffic": None, "note": note}
        tr = traffic_db.get(cls)
        if tr:   # dram__bytes_read.sum + dram__bytes_write.sum per unit from the committed `ncu --set full` capture of this workload, x this run's units per launch
            e["traffic"] = tr["dram_bytes_per_unit"] * units / max(n, 1)
            e["traffic_detail"] = dict(tr, units_per_launch=units / max(n, 1))
        return e

    trace_bytes = ray_bytes(vis_closest, cntc["closest"], 92) + 128 * cst["nodes"] + 8 * cst["prims"] + 72 * cst["tri_tests"] + 16 * cst["sphere_tests"] + (28 + 28 + 52) * cntc["closest"]
    by["trace"] = entry("k_closest_bvh + k_closest_finish + k_closest_fallback + k_wave_classify (Scene::hit: world-space BVH, the reference traversal on the winning object, full reference traversal where needed)",
                        "trace", trace_bytes, cntc["closest"], "ray",
                        "92 B ray in / hit out + 128 B per BVH node + 8 B per leaf primitive + 72 B per triangle test + 108 B of per-ray scratch and second ray read, plus the reference-order visits of the finish / fallback kernels "
                        "(64 B per object-BVH node, 96 B per instance transform, 16 B per kd node, 4 B per leaf entry, 72 B per triangle test); counts every visit, so a scene that fits the 126 MB L2 can exceed the HBM peak")
    if by["trace"]:
        by["trace"]["bvh_per_ray"] = {k: v / max(cntc["closest"], 1) for k, v in cst.items() if k != "rays"}
        by["trace"]["reference_order_visits_per_ray"] = {k: v / max(cntc["closest"], 1) for k, v in vis_closest.items()}
    by["occlude"] = entry("k_occl_bvh + k_occl_confirm + k_occl_fallback + k_shadow_apply (order-free occlusion BVH, confirmed by the reference's per-object traversal)", "occlude",
                          occl_bytes(ost, vis_occl, cntc["occlusion"]), cntc["occlusion"], "ray",
                          "60 B shadow-queue entry + 128 B per BVH node + 8 B per leaf primitive + 72 B per triangle test + 68 B per candidate + the confirming traversal's visits")
    if by["occlude"]: by["occlude"]["per_ray"] = {k: v / max(cntc["occlusion"], 1) for k, v in ost.items()}
    sst = ctx.shade_stats()      # of the counting render just above
    shade_bytes = bounces * (192 * 2 + 32 + 64) + sst["nee_bounces"] * (140 + 2 * 140) + sst["nee_terms"] * (2 * 36 + 140 + 2 * 132 + 140) + 92.0 * cntc["occlusion"]
    by["shade"] = entry("k_terminal + k_scatter<K> + k_nee_a1 + k_nee_b<K> + k_nee_a + k_nee_eval<K> (one bounce of the integrator)", "shade", shade_bytes, bounces, "bounce",
                        "SURVEY 8d's B_bounce (2 x 192 B path state + 32 B hit record + 64 B material) + the queues of the three-stage NEE: 140 B shading context written once and read by two "
                        "stages per NEE bounce (%.2f of the bounces), per NEE term 36 B of survivor queue written and read + 140 B context in the light-side stage, 132 B of term queue written and read back + 140 B context in the evaluation (%.2f terms per bounce), 92 B per queued shadow ray "
                        "(%.2f per bounce).  The class is FP64-latency and instruction-fetch bound, not HBM bound (profiles/README.md)" % (sst["nee_bounces"] / bounces, sst["nee_terms"] / bounces, n_shadow_per_bounce))
    if by["shade"]: by["shade"]["per_bounce"] = {"nee_bounces": sst["nee_bounces"] / bounces, "nee_terms": sst["nee_terms"] / bounces, "shadow_rays": n_shadow_per_bounce, "survey_8d_bytes": 480 + 128.0 * n_shadow_per_bounce}
    by = {k: v for k, v in by.items() if v}
    dominant = max(by, key=lambda k: by[k]["ms_per_step"]) if by else None
    if dominant:
        line["roofline"] = dict(by[dominant], dominant_class=dominant)
        line["roofline_by_kernel"] = by
    fp = env.get("fp64")
    if fp and "roofline" in line: line["roofline"]["fp64_peak"] = fp

    if not args.no_cpu and world == 1:
        try:
            line["cpu_baseline"] = cpu_measure(prog, integrator, spp, args.cpu_seconds if full else min(args.cpu_seconds, 6.0))
            line["vs_cpu"] = {"mrays_ratio_e2e": e2e["value"] / line["cpu_baseline"]["value"], "msamples_ratio": line["msamples_per_s"] / line["cpu_baseline"]["msamples_per_s"],
                              "rays_per_sample_gpu": line["rays_per_sample"], "rays_per_sample_cpu": line["cpu_baseline"]["rays_per_sample"],
                              "note": "same scene, same spp; rays per sample differ through the Russian-roulette threshold (per-tile pilot estimate on the GPU, per-tile running recurrence on the CPU, SURVEY A.16)"}
        except Exception as e:
            line["cpu_baseline"] = {"error": str(e)}

    if full and world == 1:
        try: line["micro"] = micro_trace(scene, dev, torch, np)
        except Exception as e: line["micro"] = {"error": str(e)}
        # film finalisation kernel (Film::rgb_image on the device): 56 B read + 3 B written per pixel, HBM bound; L2 flushed before each launch
        try:
            def film_ms(pp, sp_, h, w):
                fe = []
                for k in range(6):
                    flush.zero_(); torch.cuda.synchronize()
                    fe.append(ctx.film_encode_dev(pp, sp_, (h, w), 1.0 / max(spp, 1), 1.0, 0)[1])
                return sorted(fe[1:])[len(fe[1:]) // 2]
            fms = film_ms(px_ptr, sp_ptr, H, W)
            H4, W4 = 2160, 3840
            reps = -(-(H4 * W4) // (H * W))
            big_px = film[:W * H * 4].view(-1, 4).repeat(reps, 1)[:H4 * W4].contiguous()
            big_sp = film[W * H * 4:].view(-1, 3).repeat(reps, 1)[:H4 * W4].contiguous()
            torch.cuda.synchronize()
            fms4 = film_ms(big_px.data_ptr(), big_sp.data_ptr(), H4, W4)
            del big_px, big_sp
            line["film_encode"] = {"kernel": "k_film_encode", "unit": "GB/s", "peak": peak,
                                   "render_film": {"resolution": [W, H], "ms": fms, "bytes": W * H * 59, "achieved": W * H * 59 / fms / 1e6, "frac": W * H * 59 / fms / 1e6 / peak},
                                   "uhd_film": {"resolution": [W4, H4], "ms": fms4, "bytes": W4 * H4 * 59, "achieved": W4 * H4 * 59 / fms4 / 1e6, "frac": W4 * H4 * 59 / fms4 / 1e6 / peak},
                                   "d2h_bytes_per_pixel": 3}
        except Exception as e:
            line["film_encode"] = {"error": str(e)}
    scene.close(); del film
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="bistro", choices=list(WORKLOADS))
    ap.add_argument("--spp", type=int, default=0, help="samples per pixel per step and per GPU (0 = workload default)")
    ap.add_argument("--total-spp", type=int, default=0, help="fixed job (strong scaling): this many samples per pixel per step in total, split evenly over the GPUs")
    ap.add_argument("--wave-paths", type=int, default=0)
    ap.add_argument("--other-workloads", default="cornell,bunny,dragon,caustics_bdpt,conference", help="comma list measured after the main workload with fewer steps ('' = none); N = 1 only")
    ap.add_argument("--other-steps", type=int, default=2)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU work per cpu_baseline sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    args = ap.parse_args()
    _guard_stdout()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    from lumo_b200 import native
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: lumo_b200 has no CPU path")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    ctx = native.GpuContext(local)
    # Everything — the library's kernels, torch's copies and the NCCL reduce — is ordered on ONE non-default stream.  (The default
    # stream's handle is NULL, which lumo_gpu_ctx_set_stream reads as "use the context's own stream": the next render's memset of
    # the film would then race with the NCCL reduce / the copy of the previous step that torch ordered on its default stream.)
    side = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(side)
    assert side.cuda_stream != 0
    env = {"torch": torch, "np": np, "native": native, "dist": dist, "world": world, "rank": rank, "local": local, "dev": dev, "ctx": ctx,
           "stream": side, "flush": torch.empty(256 << 20, dtype=torch.uint8, device=dev)}
    if rank == 0:
        try: env["fp64"] = ctx.fp64_peak()
        except Exception as e: env["fp64"] = {"error": str(e)}
    line = measure(args.workload, args, env, args.steps, max(args.warmup, 3) if args.warmup >= 3 else args.warmup, full=True)
    if rank == 0 and world == 1 and args.other_workloads:
        others = {}
        for name in [n for n in args.other_workloads.split(",") if n and n != args.workload]:
            try:
                others[name] = measure(name, args, env, max(1, args.other_steps), 3, full=False)
            except Exception as e:   # a failed side measurement must not lose the main line
                others[name] = {"error": str(e)}
        line["other_workloads"] = others
    if rank == 0:
        print(json.dumps(line), file=_guard_stdout(), flush=True)
    ctx.close()
    if dist is not None:
        dist.barrier(); dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
