"""G1-G3: the CUDA traversal through the C ABI against the oracle on identical ray
batches.  Integer outputs and distances are compared bit for bit: the device code keeps the
reference's operation order and is compiled without FMA contraction."""
import numpy as np
import pytest
import oracle_lib
from conftest import small_scene, ray_batches, SMALL

pytestmark = pytest.mark.gpu
N = 20000


def _bits(a):
    return np.ascontiguousarray(a, np.float64).view(np.uint64)


@pytest.mark.parametrize("name", list(SMALL))
def test_closest_hit_ids_and_distances_bit_exact(name, gpu_ctx):
    from lumo_b200 import native
    prog, blob, _ = small_scene(name)
    O = oracle_lib.OracleScene(prog)
    G = native.GpuScene(gpu_ctx, blob)
    for kind, (o, d) in ray_batches(O, N, seed=23).items():
        eo, et, ett, eb = O.trace_closest(o, d)
        go, gt, gtt, gb = G.trace_closest(o, d)
        assert np.array_equal(eo, go), (name, kind, "object ids", int((eo != go).sum()))
        assert np.array_equal(et, gt), (name, kind, "triangle ids", int((et != gt).sum()))
        assert np.array_equal(_bits(ett), _bits(gtt)), (name, kind, "t (0 ulp)")
        assert np.array_equal(_bits(eb), _bits(gb)), (name, kind, "barycentrics (0 ulp)")
        assert (eo != 0xFFFFFFFF).sum() > 0.2 * len(eo) or kind == "incoherent"
    G.close(); O.close()


@pytest.mark.parametrize("name", list(SMALL))
def test_any_hit_and_first_found(name, gpu_ctx):
    from lumo_b200 import native
    prog, blob, _ = small_scene(name)
    O = oracle_lib.OracleScene(prog)
    G = native.GpuScene(gpu_ctx, blob)
    rs = np.random.RandomState(5)
    for kind, (o, d) in ray_batches(O, N, seed=29).items():
        _, _, t, _ = O.trace_closest(o, d)
        # t_max on both sides of the first hit, like hit_light's t_light - 1e-10
        tm = np.where(np.isfinite(t), t * rs.choice([0.5, 1.0 - 1e-12, 1.5, 3.0], size=len(t)), 10.0) - 1e-10
        assert np.array_equal(O.trace_any(o, d, tm), G.trace_any(o, d, tm)), (name, kind, "occlusion booleans")
        assert np.array_equal(_bits(O.trace_first_found(o, d)), _bits(G.trace_first_found(o, d))), (name, kind, "first-found t")
    G.close(); O.close()


def test_visit_counters_equal_oracle(gpu_ctx):
    """The N_* of the byte formula: the faithful kernel visits exactly the oracle's nodes."""
    from lumo_b200 import native
    prog, blob, _ = small_scene("bunny")
    O = oracle_lib.OracleScene(prog)
    G = native.GpuScene(gpu_ctx, blob)
    o, d = ray_batches(O, 5000, seed=31)["secondary"]
    O.counters(reset=True)
    O.trace_closest(o, d, threads=1)
    oc = O.counters(reset=True)
    gpu_ctx.closest_mode(1)                      # the device's replay of the reference traversal (the default pipeline only runs it where it has to)
    try:
        gpu_ctx.count_visits(True)
        G.trace_closest(o, d)
        gc, _ = gpu_ctx.visits()
        gpu_ctx.count_visits(False)
    finally:
        gpu_ctx.closest_mode(0)
    for k in ("tlas_nodes", "inst", "kd_nodes", "leaf_idx", "tri_tests", "sphere_tests"):
        assert oc[k] == gc[k], (k, oc[k], gc[k])


def test_edge_cases(gpu_ctx):
    from lumo_b200 import native
    prog, blob, _ = small_scene("cornell")
    G = native.GpuScene(gpu_ctx, blob)
    O = oracle_lib.OracleScene(prog)
    e = np.zeros((0, 3))
    obj, tri, t, bary = G.trace_closest(e, e)                       # empty batch
    assert len(obj) == 0
    # axis-parallel directions (zero components -> inf reciprocals), rays starting on geometry, rays leaving the box
    o = np.array([[278.0, 273.0, -800.0], [278.0, 273.0, 100.0], [278.0, 0.0, 100.0], [278.0, 273.0, 100.0], [0.0, 0.0, 0.0], [278.0, 600.0, 279.6]])
    d = np.array([[0.0, 0.0, 1.0], [0.0, -1.0, 0.0], [0.0, 1.0, 0.0], [1.0, 0.0, 0.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0]])
    eo, et, ett, eb = O.trace_closest(o, d)
    go, gt, gtt, gb = G.trace_closest(o, d)
    assert np.array_equal(eo, go) and np.array_equal(et, gt) and np.array_equal(_bits(ett), _bits(gtt))
    assert go[-1] == 0xFFFFFFFF and np.isinf(gtt[-1])
    with pytest.raises(RuntimeError):
        native.GpuScene(gpu_ctx, blob[:-16])                        # malformed blob -> status code, not a crash


@pytest.mark.parametrize("name", list(SMALL))
def test_closest_pipeline_equals_reference_traversal(name, gpu_ctx):
    """closest.cuh (world-space BVH + the reference traversal on the winning object, full reference traversal where the
    winner is not provably the reference's) against the device's own replay of the reference traversal for every ray
    (closest mode 1): the same bits, ray for ray; and only a minority of the rays needs the full replay."""
    from lumo_b200 import native
    prog, blob, _ = small_scene(name)
    O = oracle_lib.OracleScene(prog)
    G = native.GpuScene(gpu_ctx, blob)
    try:
        for kind, (o, d) in ray_batches(O, N, seed=47).items():
            gpu_ctx.closest_mode(1); so, st, stt, sb = G.trace_closest(o, d)
            gpu_ctx.closest_mode(0)
            gpu_ctx.count_visits(True); fo, ft, ftt, fb = G.trace_closest(o, d); cs = gpu_ctx.closest_stats(); gpu_ctx.count_visits(False)
            assert np.array_equal(so, fo) and np.array_equal(st, ft) and np.array_equal(_bits(stt), _bits(ftt)) and np.array_equal(_bits(sb), _bits(fb)), (name, kind)
            assert cs["rays"] == len(o) and cs["nodes"] > 0
            assert cs["fallback"] <= 0.5 * len(o), (name, kind, cs)
    finally:
        gpu_ctx.closest_mode(0); gpu_ctx.count_visits(False)
    G.close(); O.close()


def _shadow_rays(O, n, seed):
    """Shadow-ray-like batches: from surface points of the scene towards random points of its bounding box, towards points
    just in front of / just behind the first thing in that direction, and towards the surface point itself (grazing
    configurations of hit_light's `t_light - 1e-10`)."""
    rs = np.random.RandomState(seed)
    batches = ray_batches(O, n, seed)
    o1, d1 = batches["primary"]
    _, _, t, _ = O.trace_closest(o1, d1)
    ok = np.isfinite(t)
    p = o1[ok] + d1[ok] * (t[ok] * (1 - 1e-9))[:, None]                 # points on surfaces
    b = O.bounds(); lo, hi = b[:3], b[3:]
    q = lo + rs.rand(len(p), 3) * (hi - lo)
    v = q - p; dist = np.linalg.norm(v, axis=1); d = v / dist[:, None]
    out = [(p, d, dist - 1e-10)]
    _, _, t2, _ = O.trace_closest(p, d)
    f = np.isfinite(t2)
    for k in (1.0 - 1e-9, 1.0, 1.0 + 1e-9, 0.3):                          # t_max around the first blocker
        out.append((p[f], d[f], t2[f] * k - 1e-10))
    o3, d3 = batches["incoherent"]
    out.append((o3, d3, rs.rand(len(o3)) * np.linalg.norm(hi - lo)))
    return out


@pytest.mark.parametrize("name", list(SMALL))
def test_occlusion_bvh_equals_reference_traversal_and_oracle(name, gpu_ctx):
    """G2: the order-free occlusion BVH (+ confirmation by the reference's per-object traversal, occlude.cuh) returns the
    oracle's booleans, and the same booleans as the device's own replay of the reference traversal (occlusion mode 1)."""
    from lumo_b200 import native
    prog, blob, _ = small_scene(name)
    O = oracle_lib.OracleScene(prog)
    G = native.GpuScene(gpu_ctx, blob)
    n_rays = 0
    try:
        for o, d, tm in _shadow_rays(O, N, seed=43):
            want = O.trace_any(o, d, tm)
            gpu_ctx.occlusion_mode(0); fast = G.trace_any(o, d, tm)
            gpu_ctx.occlusion_mode(1); slow = G.trace_any(o, d, tm)
            assert np.array_equal(slow, want), (name, "reference traversal on the device vs oracle", int((slow != want).sum()))
            assert np.array_equal(fast, want), (name, "occlusion BVH vs oracle", int((fast != want).sum()))
            n_rays += len(want)
    finally:
        gpu_ctx.occlusion_mode(0)
    assert n_rays > 2 * N
    G.close(); O.close()


def test_occlusion_bvh_counters_and_render_cross_check(gpu_ctx):
    """Mode 2 runs both kernels on every shadow ray of a render: zero disagreements, and the film equals the default mode's
    up to the order of the atomic adds.  The counting instantiation reports the BVH work (bench.py's roofline input)."""
    from lumo_b200 import native
    for name in ("bunny", "bistro", "conference", "caustics"):
        prog, blob, ig = small_scene(name)
        G = native.GpuScene(gpu_ctx, blob)
        try:
            gpu_ctx.count_visits(False)
            gpu_ctx.occlusion_mode(0); a = G.render(integrator=0, spp=16, seed=9)
            gpu_ctx.occlusion_mode(2); b = G.render(integrator=0, spp=16, seed=9)
            st = gpu_ctx.occlusion_stats()
            gpu_ctx.occlusion_mode(1); c = G.render(integrator=0, spp=16, seed=9)
        finally:
            gpu_ctx.occlusion_mode(0)
        assert st["mismatches"] == 0, (name, st)
        assert b[2]["occlusion_mismatches"] == 0, (name, b[2])      # also counts a light test passing outside the light's bounding sphere (k_nee_b)
        assert a[2]["occlusion"] == b[2]["occlusion"] == c[2]["occlusion"] > 0 and a[2]["closest"] == c[2]["closest"]
        assert np.allclose(a[0], c[0], rtol=1e-11, atol=1e-13) and np.allclose(a[0], b[0], rtol=1e-11, atol=1e-13), name
        gpu_ctx.count_visits(True)
        G.render(integrator=0, spp=2, seed=9)
        st = gpu_ctx.occlusion_stats(); gpu_ctx.count_visits(False)
        assert st["nodes"] > 0 and st["tri_tests"] + st["sphere_tests"] > 0 and st["candidates"] == st["confirmed"] + st["fallback"], (name, st)
        assert st["fallback"] <= 1e-3 * max(st["candidates"], 1) + 2, (name, st)
        G.close()


def test_special_rays_bit_exact(gpu_ctx):
    """Edge cases of the arithmetic: zero direction components (inf reciprocals, 0*inf = NaN dropped by
    min/max), negative zero, origins exactly on box planes / kd split planes / triangle planes, finite t_max on both sides of
    the hit, directions that are not normalised, NaN and infinite inputs.  GPU and oracle must agree bit for bit — including
    on what they return for garbage."""
    from lumo_b200 import native
    for name in ("cornell", "bunny", "conference"):
        prog, blob, _ = small_scene(name)
        O = oracle_lib.OracleScene(prog); G = native.GpuScene(gpu_ctx, blob)
        B = G.blob
        rs = np.random.RandomState(3)
        lo, hi = np.array(B.params["bounds_lo"]), np.array(B.params["bounds_hi"])
        n = 4000
        o = lo + rs.rand(n, 3) * (hi - lo)
        d = rs.randn(n, 3); d /= np.linalg.norm(d, axis=1, keepdims=True)
        # axis-parallel and plane-parallel directions, with both signs of zero
        d[0:600, 0] = 0.0; d[200:800, 1] = -0.0; d[400:1000, 2] = 0.0
        d[:1000] /= np.maximum(np.linalg.norm(d[:1000], axis=1, keepdims=True), 1e-300)
        # origins on kd split planes and on triangle vertices of the first tree
        T = B.kd_trees[0]
        nodes = B.kd_nodes[int(T["root"]):int(T["root"]) + 64]
        inner = nodes[(nodes["b"] & 0x80000000) == 0]
        for k, nd in enumerate(inner[:200]):
            o[1000 + k, int(nd["b"])] = float(nd["point"])
        tv = B.tri_verts[int(T["tri_base"]):int(T["tri_base"]) + 200]
        o[1200:1200 + len(tv)] = tv["a"]
        # origins on the scene bounds
        o[1500:1700, 0] = lo[0]; o[1700:1900, 1] = hi[1]
        # un-normalised directions (instance-local rays are like this inside the traversal; the API takes them as given)
        d[2000:2400] *= rs.rand(400, 1) * 10 + 0.1
        # garbage
        d[2400:2420] = 0.0; d[2420:2440, 0] = np.nan; o[2440:2460, 1] = np.inf; d[2460:2480] = np.inf; o[2480:2500] = 1e300
        for oo, dd in ((o, d),):
            with np.errstate(all="ignore"):
                eo, et, ett, eb = O.trace_closest(oo, dd)
                go, gt, gtt, gb = G.trace_closest(oo, dd)
            assert np.array_equal(eo, go), (name, np.nonzero(eo != go)[0][:10])
            assert np.array_equal(et, gt), name
            assert np.array_equal(_bits(ett), _bits(gtt)) and np.array_equal(_bits(eb), _bits(gb)), name
            ef = O.trace_first_found(oo, dd); gf = G.trace_first_found(oo, dd)
            assert np.array_equal(_bits(ef), _bits(gf)), name
            for scale in (0.25, 1.0, 1.0 + 1e-15, 4.0):
                tm = np.where(np.isfinite(ett), ett * scale, 1.0)
                assert np.array_equal(O.trace_any(oo, dd, tm), G.trace_any(oo, dd, tm)), (name, scale)
        # finite t_max on the closest-hit entry point (the reference always passes +inf; the C ABI allows it)
        tm = np.where(np.isfinite(ett), ett * rs.choice([0.5, 1.0, 2.0], size=n), 3.0)
        go2, gt2, gtt2, _ = G.trace_closest(o, d, t_max=tm)
        keep = np.isfinite(ett) & (tm >= ett * 1.5)
        assert np.array_equal(go2[keep], go[keep]) and np.array_equal(_bits(gtt2[keep]), _bits(gtt[keep]))
        G.close(); O.close()
