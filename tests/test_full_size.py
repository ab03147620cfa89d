"""BASELINE-size scenes (full bunny and bistro stand-ins, full resolution) through properties that do not need
the oracle to finish a full-size run: object contract (hit <=> first-found finite, test_util.rs:26-45),
occlusion consistent with the closest hit, sample ranges composing, determinism, exact filter-weight sums,
and a bounded-sample oracle cross-check of hit ids on the full-size structure."""
import numpy as np
import pytest
import oracle_lib

pytestmark = pytest.mark.gpu
_cache = {}


def _full(name):
    if name not in _cache:
        import bench
        _cache[name] = bench.build_workload(name)
    return _cache[name]


def _bits(a):
    return np.ascontiguousarray(a, np.float64).view(np.uint64)


FULL = ["cornell", "bunny", "dragon", "caustics_bdpt", "conference", "bistro"]      # the BASELINE configs at their full sizes


def _device_rays(B, n, seed, torch_dev="cuda:0"):
    """n incoherent rays inside the scene bounds (origins uniform, directions uniform on the sphere), generated with numpy."""
    rs = np.random.RandomState(seed)
    lo, hi = np.array(B.params["bounds_lo"]), np.array(B.params["bounds_hi"])
    o = lo + rs.rand(n, 3) * (hi - lo)
    z = 1 - 2 * rs.rand(n); ph = 2 * np.pi * rs.rand(n); r = np.sqrt(np.maximum(1 - z * z, 0))
    return o, np.stack([r * np.cos(ph), r * np.sin(ph), z], -1)


@pytest.mark.parametrize("name", FULL)
def test_full_size_ten_million_rays_both_pipelines(name, gpu_ctx):
    """G1 / G2 at full size, >= 10^7 rays per config: the default pipelines (world-space BVH + the reference traversal where
    it decides: closest.cuh, occlude.cuh) against the device's replay of the reference traversal for every ray — ids,
    distances, barycentrics and occlusion booleans bit for bit; then a bounded sample of each batch kind (primary rays of
    the config's camera, incoherent rays, bounce rays off the first hit) against the oracle on the same full-size scene."""
    from lumo_b200 import native
    from conftest import ray_batches
    prog, blob, integrator, _ = _full(name)
    G = native.GpuScene(gpu_ctx, blob)
    n_total = 0
    try:
        for chunk in range(5):
            o, d = _device_rays(G.blob, 2_100_000, 100 + chunk)
            gpu_ctx.closest_mode(0); fo, ft, ftt, fb = G.trace_closest(o, d)
            gpu_ctx.closest_mode(1); so, st, stt, sb = G.trace_closest(o, d)
            assert np.array_equal(fo, so) and np.array_equal(ft, st) and np.array_equal(_bits(ftt), _bits(stt)) and np.array_equal(_bits(fb), _bits(sb)), (name, chunk)
            # shadow-ray-like segments: up to just before / just behind the first hit, and random lengths
            rs = np.random.RandomState(200 + chunk)
            tm = np.where(np.isfinite(ftt), ftt * rs.choice([0.5, 1.0 - 1e-9, 1.0 + 1e-9, 2.0], size=len(ftt)), rs.rand(len(ftt)) * 100.0) - 1e-10
            gpu_ctx.occlusion_mode(0); fa = G.trace_any(o, d, tm)
            gpu_ctx.occlusion_mode(1); sa = G.trace_any(o, d, tm)
            assert np.array_equal(fa, sa), (name, chunk, int((fa != sa).sum()))
            assert 0.02 < fa.mean() < 0.98 or name == "cornell"
            n_total += len(o)
    finally:
        gpu_ctx.closest_mode(0); gpu_ctx.occlusion_mode(0)
    assert n_total >= 10_000_000
    O = oracle_lib.OracleScene(prog)
    for kind, (o, d) in ray_batches(O, 20000, seed=53).items():
        eo, et, ett, eb = O.trace_closest(o, d)
        go, gt, gtt, gb = G.trace_closest(o, d)
        assert np.array_equal(eo, go) and np.array_equal(et, gt) and np.array_equal(_bits(ett), _bits(gtt)) and np.array_equal(_bits(eb), _bits(gb)), (name, kind)
        tm = np.where(np.isfinite(ett), ett * 0.999, 5.0)
        assert np.array_equal(O.trace_any(o, d, tm), G.trace_any(o, d, tm)), (name, kind)
    O.close(); G.close()


@pytest.mark.parametrize("name", FULL)
def test_full_size_render_cross_check(name, gpu_ctx):
    """Every shadow ray of a full-size render through both occlusion paths (mode 2): zero disagreements; and the film and the
    counters of the default pipelines equal those of the all-reference-traversal render (closest mode 1 + occlusion mode 1)."""
    from lumo_b200 import native
    prog, blob, integrator, _ = _full(name)
    if integrator == 2: integrator = 0                 # the shadow queue belongs to PathTrace / DirectLight
    G = native.GpuScene(gpu_ctx, blob)
    spp = 1 if name == "bistro" else 4
    try:
        gpu_ctx.occlusion_mode(2); a = G.render(integrator=integrator, spp=spp, seed=5, rr_delta=0.05)
        gpu_ctx.occlusion_mode(1); gpu_ctx.closest_mode(1); b = G.render(integrator=integrator, spp=spp, seed=5, rr_delta=0.05)
    finally:
        gpu_ctx.occlusion_mode(0); gpu_ctx.closest_mode(0)
    assert a[2]["occlusion_mismatches"] == 0, (name, a[2])
    for k in ("camera_paths", "closest", "occlusion", "cost", "max_depth", "nonfinite"):
        assert a[2][k] == b[2][k], (name, k, a[2][k], b[2][k])
    assert a[2]["occlusion"] > 200_000
    scale = np.abs(b[0]).max()
    assert np.allclose(a[0], b[0], rtol=1e-10, atol=1e-13 * scale), name
    G.close()


@pytest.mark.parametrize("name", ["bunny", "bistro"])
def test_full_size_ray_properties(name, gpu_ctx):
    from lumo_b200 import native
    prog, blob, integrator, _ = _full(name)
    G = native.GpuScene(gpu_ctx, blob)
    B = G.blob
    rs = np.random.RandomState(41)
    n = 400000
    lo, hi = np.array(B.params["bounds_lo"]), np.array(B.params["bounds_hi"])
    o = lo + rs.rand(n, 3) * (hi - lo)
    z = 1 - 2 * rs.rand(n); ph = 2 * np.pi * rs.rand(n); r = np.sqrt(np.maximum(1 - z * z, 0))
    d = np.stack([r * np.cos(ph), r * np.sin(ph), z], -1)
    obj, tri, t, bary = G.trace_closest(o, d)
    hit = obj != 0xFFFFFFFF
    assert hit.mean() > 0.5
    assert np.isfinite(t[hit]).all() and (t[hit] >= 0).all() and np.isinf(t[~hit]).all()
    n_obj = int(B.params["n_objects"]) + int(B.params["n_lights"])
    assert (obj[hit] < n_obj).all()
    kd = (B.objects["kind"][obj[hit]] <= 1)
    ntri = B.kd_trees["n_tris"][B.objects["geom"][obj[hit]][kd]]
    assert (tri[hit][kd] < ntri).all()
    bsum = bary[hit][kd].sum(axis=1)
    assert (bary[hit][kd] >= -1e-9).all() and (bsum <= 1 + 1e-9).all()
    # hit is Some  =>  hit_t is finite (the converse can fail: kdtree.rs:165 / SURVEY A.6)
    tf = G.trace_first_found(o, d)
    assert np.isfinite(tf[hit]).all()
    assert (np.isfinite(tf) & ~hit).mean() < 1e-3
    # something is hit before t*(1+1e-9): the occlusion query must say so; nothing can be hit on a miss of both queries
    assert G.trace_any(o[hit], d[hit], t[hit] * (1 + 1e-9) + 1e-9).all()
    both_miss = ~hit & np.isinf(tf)
    if both_miss.any():
        assert not G.trace_any(o[both_miss], d[both_miss], np.full(int(both_miss.sum()), 1e30)).any()
    # a bounded sample against the oracle on the full-size structure: ids and distances bit for bit
    O = oracle_lib.OracleScene(prog)
    k = 20000
    eo, et, ett, eb = O.trace_closest(o[:k], d[:k])
    assert np.array_equal(eo, obj[:k]) and np.array_equal(et, tri[:k]) and np.array_equal(_bits(ett), _bits(t[:k])) and np.array_equal(_bits(eb), _bits(bary[:k]))
    O.close(); G.close()


def test_full_size_render_properties(gpu_ctx):
    """Full-resolution bunny: ranges compose, a render is reproducible, and the film's weight plane is exactly what the
    filter and the sample positions give (no sample lost or duplicated by the queues)."""
    from lumo_b200 import native
    prog, blob, integrator, _ = _full("bunny")
    G = native.GpuScene(gpu_ctx, blob)
    full, _, cf, _, _ = G.render(integrator=integrator, spp=4, seed=9, rr_delta=0.05)
    a, _, ca, _, _ = G.render(integrator=integrator, spp=4, seed=9, rr_delta=0.05, spp_begin=0, spp_end=1)
    b, _, cb, _, _ = G.render(integrator=integrator, spp=4, seed=9, rr_delta=0.05, spp_begin=1, spp_end=4)
    assert cf["camera_paths"] == 4 * G.res_x * G.res_y == ca["camera_paths"] + cb["camera_paths"]
    assert ca["closest"] + cb["closest"] == cf["closest"] and ca["occlusion"] + cb["occlusion"] == cf["occlusion"] and ca["cost"] + cb["cost"] == cf["cost"]
    # f64 atomics accumulate in a different order; colour channels can cancel (negative lobes of XYZ -> RGB)
    assert np.allclose(a + b, full, rtol=1e-9, atol=1e-12)
    again, _, cg, _, _ = G.render(integrator=integrator, spp=4, seed=9, rr_delta=0.05, wave_paths=300000)
    assert cg == {**cf, "gpu_launches": cg["gpu_launches"], "iterations": cg["iterations"]}
    assert np.allclose(again, full, rtol=1e-9, atol=1e-12)
    assert cf["nonfinite"] == 0 and np.isfinite(full).all()
    assert (full[..., 3] > 0).all() and (full[..., :3] >= -1e-9).sum() >= 0
    G.close()
