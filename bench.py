#!/usr/bin/env python
"""bench.py — Mrays/s and Msamples/s of the lumo hot path on B200 (BASELINE.json metric).

A "step" is one pass of the hot path over the workload: `--spp` samples per pixel of the whole
image through the wavefront pipeline (camera rays -> closest-hit traversal -> shading/NEE ->
occlusion traversal -> film).  rays = closest-hit queries + occlusion queries, counted on the
device.  With N GPUs (torchrun) every rank renders its own sample range of a N*spp render with the
scene replicated (weak scaling) and the film accumulators are combined with ONE NCCL reduce
inside the timed region.

  value     whole-job Mrays/s, scene resident in HBM, film left on the device
  e2e       the same metric through the public C-ABI calls with HOST buffers: scene blob upload
            (H2D) + render + film download (D2H) inside the timed region
  roofline  the closest-hit trace kernel: algorithmic bytes (DESIGN.md byte formula x visit
            counters from an untimed counting pass of the same workload) / its CUDA-event time
  cpu_baseline  the C++ restatement of lumo's CPU renderer (oracle/, reference schedule, all host
            cores) on a bounded sample of the same workload
`--impl reference` runs only that CPU restatement (the reference is Rust and cannot be built here)."""
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
    "bunny": (dict(), 0, 32),
    "cornell": (dict(), 0, 64),
    "dragon": (dict(), 0, 8),
    "caustics_bdpt": (dict(), 2, 2),
    "conference": (dict(), 0, 16),
    "conference_dl": (dict(), 1, 32),
    "bistro": (dict(), 0, 2),
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


def build_workload(name):
    from lumo_b200 import scenes, native
    kw, integrator, spp = WORKLOADS[name]
    s, cam, _ = scenes.CONFIGS[name.split("_")[0]](**kw)
    prog = s._program(cam)
    blob = native.build_blob(prog)
    return prog, blob, integrator, spp


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
    """DESIGN.md byte formula: per-ray queue traffic + nodes / instances / leaf entries / triangles visited."""
    return fixed * n_rays + 64 * v["tlas_nodes"] + 96 * v["inst"] + 16 * v["kd_nodes"] + 4 * v["leaf_idx"] + 72 * v["tri_tests"] + 16 * v["sphere_tests"]


def micro_trace(scene, dev, torch, np, n=1 << 22):
    """lumo_gpu_trace_closest_dev on device-resident batches; kernel time from the library's CUDA events."""
    P = scene.blob.params; cam = P["camera"]
    W, H = int(cam["res_x"]), int(cam["res_y"])
    g = torch.Generator(device=dev); g.manual_seed(23)
    # primary rays: Camera::generate_ray for a pinhole camera (camera.rs:257-268), evaluated in f64 on the host side of the bench
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


def cpu_baseline(prog, integrator, threads, spp=1, total_spp=None, seed=1):
    """The C++ restatement of lumo's CPU renderer (reference schedule: 16x16 tiles x 256-sample batches from a
    shared queue, per-tile xorshift; renderer.rs:174-204), -O3 -march=native, `threads` workers."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    O = oracle_lib.OracleScene(prog, native=True)
    t0 = time.time()
    _, _, cnt, _ = O.render(integrator=integrator, spp=spp, seed=seed, rng_mode=0, threads=threads)
    dt = time.time() - t0
    O.close()
    rays = cnt["closest"] + cnt["occlusion"]
    return rays / dt / 1e6, cnt["camera_paths"] / dt / 1e6, dt, cnt


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
    prog, blob, integrator, spp_default = build_workload(args.workload)
    threads = os.cpu_count() or 1
    vals, samps, dts = [], [], []
    for i in range(args.warmup + args.steps):
        v, s, dt, cnt = cpu_baseline(prog, integrator, threads, spp=1, seed=1 + i)
        if i >= args.warmup:
            vals.append(v); samps.append(s); dts.append(dt)
    val = sum(vals) / len(vals)
    line = {"impl": "reference", "metric": "Mrays/s", "value": val, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * sum(dts) / len(dts), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "description": DESCR[args.workload], "integrator": ["PathTrace", "DirectLight", "BDPathTrace"][integrator]},
            "msamples_per_s": sum(samps) / len(samps),
            "cpu_baseline": {"value": val, "unit": "Mrays/s", "cores": threads, "kind": "port",
                             "sample": "1 spp of the full-resolution workload per step (lumo CPU path, C++ restatement; the Rust reference cannot be built offline)"},
            "e2e": {"value": val, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), file=_guard_stdout(), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="bunny", choices=list(WORKLOADS))
    ap.add_argument("--spp", type=int, default=0, help="samples per pixel per step and per GPU (0 = workload default)")
    ap.add_argument("--wave-paths", type=int, default=0)
    ap.add_argument("--other-scenes", default="cornell,dragon,caustics_bdpt,conference,conference_dl,bistro,textured", help="comma list measured briefly after the main workload ('' = none)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
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

    prog, blob, integrator, spp = build_workload(args.workload)
    if args.spp: spp = args.spp
    total_spp = spp * world
    s0, s1 = rank * spp, (rank + 1) * spp
    ctx = native.GpuContext(local)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    scene = native.GpuScene(ctx, blob)
    W, H = scene.res_x, scene.res_y
    film = torch.zeros(W * H * 7, dtype=torch.float64, device=dev)      # pixels[W*H*4] then splats[W*H*3]
    px_ptr, sp_ptr = film.data_ptr(), film.data_ptr() + W * H * 4 * 8
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2

    def step(seed):
        cnt, ms = scene.render_dev(px_ptr, sp_ptr, integrator=integrator, seed=seed, spp_begin=s0, spp_end=s1, total_spp=total_spp, wave_paths=args.wave_paths)
        if dist is not None:
            dist.reduce(film, dst=0)       # the one collective of the path: film accumulators over NVLink
        return cnt

    def barrier():
        if dist is not None: dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(100 + i)
    barrier()
    clocks = ClockSampler(local) if rank == 0 else None
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    tot = {"closest": 0, "occlusion": 0, "camera_paths": 0, "cost": 0, "gpu_launches": 0, "iterations": 0}
    ktimes = {"regen": [0.0, 0], "trace": [0.0, 0], "shade": [0.0, 0], "occlude": [0.0, 0]}
    seeds = [1 + i for i in range(args.steps)]
    barrier()
    for i in range(args.steps):
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

    e2e_multi = None
    if dist is not None:
        # multi-GPU e2e (all ranks take part in the reduce): render + NCCL reduce + film download on rank 0, wall clock
        barrier(); t0 = time.perf_counter()
        cnt2 = step(seeds[-1])
        if rank == 0: host_film = film.cpu()
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        r2 = torch.tensor([cnt2["closest"] + cnt2["occlusion"]], dtype=torch.float64, device=dev)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX); dist.all_reduce(r2, op=dist.ReduceOp.SUM)
        e2e_multi = float(r2.item()) / float(dt.item()) / 1e6
    if rank != 0:
        if dist is not None:
            dist.barrier(); dist.destroy_process_group()
        return 0

    # ---- rank 0 only from here: counting pass, e2e, cpu baseline (N=1), other scenes ------------
    line = {"metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "description": DESCR[args.workload], "integrator": ["PathTrace", "DirectLight", "BDPathTrace"][integrator],
                       "resolution": [W, H], "spp_per_step_per_gpu": spp, "parallelism": "scene replicated, samples sharded, 1 NCCL reduce" if world > 1 else "single GPU",
                       "l2": "256 MiB buffer written between timed steps (L2 flush); wave state (2^21 slots, ~0.7 GB) and film (44 MB) also exceed L2"},
            "msamples_per_s": paths / (ms_total * 1e-3) / 1e6,
            "rays": {"closest_hit": closest, "occlusion": occl, "camera_paths": paths, "reference_style_cost": tot["cost"]},
            "gpu_launches": int(launches), "clocks": clk}
    line["kernel_ms_per_step"] = {k: v[0] / args.steps for k, v in ktimes.items()}

    # roofline of the dominant kernel (closest-hit trace): one untimed counting pass of the last timed step
    ctx.count_visits(True)
    cntc, _ = scene.render_dev(px_ptr, sp_ptr, integrator=integrator, seed=seeds[-1], spp_begin=s0, spp_end=s1, total_spp=total_spp, wave_paths=args.wave_paths)
    vis_closest, vis_occl = ctx.visits()
    occl_stats = ctx.occlusion_stats()
    ctx.count_visits(False)
    # wave trace kernel per ray: 4 B slot index + 48 B ray read, 40 B hit record written
    bytes_trace = ray_bytes(vis_closest, cntc["closest"], 92)
    bytes_occl = ray_bytes(vis_occl, cntc["occlusion"], 60)
    trace_ms, trace_n = ktimes["trace"][0] / args.steps, ktimes["trace"][1] / args.steps
    occl_ms = ktimes["occlude"][0] / args.steps
    peaks = {}
    try: peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception: pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = bytes_trace / (trace_ms * 1e-3) / 1e9
    line["roofline"] = {"bound": "hbm", "kernel": "k_wave_trace (closest-hit)", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                        "algorithmic_bytes_per_launch": bytes_trace / max(trace_n, 1), "launches_per_step": trace_n, "avg_launch_ms": trace_ms / max(trace_n, 1),
                        "bytes_per_ray": bytes_trace / max(cntc["closest"], 1), "visits_per_ray": {k: v / max(cntc["closest"], 1) for k, v in vis_closest.items()},
                        "traffic": None, "occlusion_bvh_per_ray": {k: v / max(cntc["occlusion"], 1) for k, v in occl_stats.items()},
                        "occlusion_kernel": {"achieved": bytes_occl / (occl_ms * 1e-3) / 1e9 if occl_ms > 0 else None, "bytes_per_ray": bytes_occl / max(cntc["occlusion"], 1)},
                        "note": "algorithmic bytes count every node/triangle visit; a scene that fits the 126 MB L2 is served from L2/L1, so this can exceed the HBM peak (see profiles/ for dram__bytes)"}
    try:   # dram__bytes_read.sum + dram__bytes_write.sum of that kernel from the committed `ncu --set full` capture, per launch like `achieved`
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.workload)
        if tr:
            rays_per_launch = cntc["closest"] / max(trace_n, 1)
            line["roofline"]["traffic"] = tr["dram_bytes_per_ray"] * rays_per_launch
            line["roofline"]["traffic_detail"] = dict(tr, rays_per_launch=rays_per_launch, note="measured DRAM bytes per ray x this run's rays per launch")
    except Exception:
        pass

    # e2e: public C-ABI calls with host buffers — upload the blob, render, download the film
    if world == 1:
        ctx.set_stream(None)
        # host side of the call: the blob and the film buffers live in pinned host memory, as a caller that renders many frames keeps them
        pin_blob = torch.frombuffer(bytearray(blob), dtype=torch.uint8).pin_memory()
        pin_px = torch.empty((H, W, 4), dtype=torch.float64).pin_memory(); pin_sp = torch.empty((H, W, 3), dtype=torch.float64).pin_memory()
        e_rays, e_t = 0.0, 0.0
        for i in range(1 + min(args.steps, 5)):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            sc2 = native.GpuScene(ctx, blob, host_ptr=pin_blob.data_ptr())
            t1 = time.perf_counter()
            px, sp, cnt2, _, dev_ms = sc2.render(integrator=integrator, spp=spp, seed=seeds[min(i, len(seeds) - 1)], wave_paths=args.wave_paths, pixels=pin_px.numpy(), splats=pin_sp.numpy())
            t2 = time.perf_counter()
            sc2.close()
            dt = time.perf_counter() - t0
            print("e2e step %d: %.1f ms wall (upload %.1f, render call %.1f of which device %.1f, iterations %d)" % (i, 1e3 * dt, 1e3 * (t1 - t0), 1e3 * (t2 - t1), dev_ms, cnt2["iterations"]), file=sys.stderr)
            if i > 0: e_rays += cnt2["closest"] + cnt2["occlusion"]; e_t += dt
        line["e2e"] = {"value": e_rays / e_t / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": len(blob) + 64, "d2h_bytes_per_step": W * H * 56 + 64,
                       "what": "lumo_gpu_scene_upload (pinned host blob) + lumo_gpu_render (pinned host film buffers) + lumo_gpu_scene_destroy per step, wall clock, mean of up to 5 steps after one discarded"}
        ctx.set_stream(stream.cuda_stream)
    else:
        line["e2e"] = {"value": e2e_multi, "unit": "Mrays/s", "h2d_bytes_per_step": 128, "d2h_bytes_per_step": W * H * 56,
                       "what": "render_dev on every rank + NCCL reduce + film download on rank 0, wall clock (max over ranks); scene already resident"}

    # micro: the closest-hit kernel alone on caller-supplied batches (SURVEY 8d "micro"): 2^22 primary rays of the config
    # camera and 2^22 incoherent rays (origins uniform in the scene bounds, directions uniform on the sphere)
    if world == 1:
        try:
            line["micro"] = micro_trace(scene, dev, torch, np)
        except Exception as e:
            line["micro"] = {"error": str(e)}

    # film finalisation kernel (Film::rgb_image on the device): 56 B read + 3 B written per pixel, HBM bound; L2 flushed
    # before each launch.  (i) the accumulators the timed steps left in HBM, (ii) a 3840x2160 film of the same values
    # tiled (464 MB in, larger than L2): at (i)'s size the launch's fixed cost decides the time, (ii) shows the bandwidth
    if world == 1:
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

    if world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        v, s, dt, ccnt = cpu_baseline(prog, integrator, threads, spp=1)
        line["cpu_baseline"] = {"value": v, "unit": "Mrays/s", "cores": threads, "kind": "port", "msamples_per_s": s, "seconds": dt,
                                "sample": "1 spp of the same full-resolution workload (lumo CPU path, C++ restatement with the reference tile/batch schedule; the Rust reference cannot be built offline)"}

    if world == 1 and args.other_scenes:
        others = {}
        for name in [n for n in args.other_scenes.split(",") if n and n != args.workload]:
            try:
                p2, b2, ig2, spp2 = build_workload(name)
                sc2 = native.GpuScene(ctx, b2)
                f2 = torch.zeros(sc2.res_x * sc2.res_y * 7, dtype=torch.float64, device=dev)
                for k in range(2):
                    flush.zero_()
                    cnt2, ms2 = sc2.render_dev(f2.data_ptr(), f2.data_ptr() + sc2.res_x * sc2.res_y * 32, integrator=ig2, seed=5 + k, spp=spp2)
                kt = ctx.kernel_times()
                others[name] = {"mrays_per_s": (cnt2["closest"] + cnt2["occlusion"]) / ms2 / 1e3, "msamples_per_s": cnt2["camera_paths"] / ms2 / 1e3, "ms": ms2, "spp": spp2,
                                "resolution": [sc2.res_x, sc2.res_y], "kernel_ms": {k: v[0] for k, v in kt.items()}, "description": DESCR[name]}
                sc2.close(); del f2
            except Exception as e:   # a failed side measurement must not lose the main line
                others[name] = {"error": str(e)}
        line["other_scenes"] = others

    print(json.dumps(line), file=_guard_stdout(), flush=True)
    scene.close(); ctx.close()
    if dist is not None:
        dist.barrier(); dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
