"""G4: the wavefront integrators against the oracle run with the SAME counter-based random streams
(oracle rng_mode=1).  Same draws + same operation order + the same transcendental functions
(csrc/common/lumo_math.h is compiled into the device kernels, the host builder and the oracle alike) make
every GPU path the oracle's path, decision for decision.  Round 1 — CUDA's libm on one side, glibc's on the
other — lost 0.3-10 % of the pixels to last-bit differences in sampled directions flipping geometric ties;
with the shared functions NO pixel differs (profiles/r2_parity_same_streams.txt: 21 scene x integrator
combinations, 0 pixels beyond 1e-12).  So: (1) per-sample comparison — every pixel equal to the rounding of
the film's atomic adds, closest-hit and cost counters equal exactly, RR thresholds equal to rounding;
(2) relMSE against the oracle's reference schedule no larger than the oracle's own seed-to-seed relMSE."""
import numpy as np
import pytest
import oracle_lib
from conftest import small_scene

pytestmark = pytest.mark.gpu


def _rgb(px):
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.nan_to_num(px[..., :3] / px[..., 3:4])


def _relmse(a, b):
    return float(np.mean((a - b) ** 2 / (b ** 2 + 1e-3)))


CASES = [("cornell", 0, 2), ("cornell", 1, 2), ("bunny", 0, 2), ("dragon", 0, 2), ("conference", 0, 2), ("conference", 1, 2), ("bistro", 0, 1), ("bistro", 1, 1), ("caustics", 0, 2),
         ("textured", 0, 2), ("textured", 1, 2)]      # every Texture kind, bump maps, image light + environment (scenes.textured)


@pytest.mark.parametrize("name,integrator,spp", CASES)
def test_per_sample_parity_same_streams(name, integrator, spp, gpu_ctx):
    from lumo_b200 import native
    prog, blob, _ = small_scene(name, box_filter=True)
    O = oracle_lib.OracleScene(prog)
    G = native.GpuScene(gpu_ctx, blob)
    epx, esp, ecnt, edel = O.render(integrator=integrator, spp=spp, seed=7, rng_mode=1)
    gpx, gsp, gcnt, gdel, ms = G.render(integrator=integrator, spp=spp, seed=7)
    assert gcnt["camera_paths"] == ecnt["camera_paths"] == spp * O.res_x * O.res_y
    assert gcnt["nonfinite"] == 0
    # filter weights do not depend on shading: equal up to summation order
    assert np.allclose(gpx[..., 3], epx[..., 3], rtol=1e-12, atol=0)
    # per-tile RR thresholds from the pilot passes: a tile whose 128 pilot paths all agree matches to rounding
    rel = np.abs(gdel / edel - 1)
    assert rel.max() < 1e-12, (np.median(rel), rel.max())                      # the pilot paths are the same paths: sums differ by rounding only
    # the same paths: the same number of closest-hit queries and the same reference-style cost, exactly
    assert gcnt["closest"] == ecnt["closest"] and gcnt["cost"] == ecnt["cost"], (gcnt, ecnt)
    assert gcnt["occlusion"] <= ecnt["occlusion"]                              # shadow rays whose MIS contribution is exactly zero are not queued on the device
    e, g = _rgb(epx), _rgb(gpx)
    bad = (np.abs(g - e) > 1e-11 * (np.abs(e) + 1e-6)).any(axis=-1)
    assert not bad.any(), ("pixels whose samples differ beyond the rounding of the film's atomic adds", int(bad.sum()), float(np.abs(g - e).max()))
    G.close(); O.close()


@pytest.mark.parametrize("name,spp", [("cornell", 2), ("caustics", 2), ("bunny", 1), ("textured", 1)])
def test_bdpt_per_sample_parity_same_streams(name, spp, gpu_ctx):
    """BDPathTrace (bd_path_trace.rs:23-75): main samples and light-tracing splats against the oracle with
    shared Philox streams.  Splats land on other pixels, so they are compared as a film: pixels where both
    agree to 1e-9, plus totals."""
    from lumo_b200 import native
    prog, blob, _ = small_scene(name, box_filter=True)
    O = oracle_lib.OracleScene(prog)
    G = native.GpuScene(gpu_ctx, blob)
    epx, esp, ecnt, edel = O.render(integrator=2, spp=spp, seed=7, rng_mode=1)
    gpx, gsp, gcnt, gdel, ms = G.render(integrator=2, spp=spp, seed=7)
    assert gcnt["camera_paths"] == ecnt["camera_paths"] == spp * O.res_x * O.res_y
    assert gcnt["nonfinite"] == 0                      # also: no subpath was cut at the device vertex cap (high half)
    assert np.allclose(gpx[..., 3], epx[..., 3], rtol=1e-12, atol=0)
    rel = np.abs(gdel / edel - 1)
    assert rel.max() < 1e-12, (np.median(rel), rel.max())
    assert gcnt["closest"] == ecnt["closest"] and gcnt["occlusion"] == ecnt["occlusion"] and gcnt["cost"] == ecnt["cost"], (gcnt, ecnt)
    e, g = _rgb(epx), _rgb(gpx)
    bad = (np.abs(g - e) > 1e-9 * (np.abs(e) + 1e-6)).any(axis=-1)    # a sample's radiance is a sum over its (s, t) terms, added with atomics in any order
    assert not bad.any(), ("pixels whose main samples differ beyond rounding", int(bad.sum()))
    sbad = (np.abs(gsp - esp) > 1e-9 * (np.abs(esp) + 1e-6)).any(axis=-1)
    assert not sbad.any(), ("pixels whose splats differ beyond rounding", int(sbad.sum()))
    G.close(); O.close()


@pytest.mark.parametrize("name,integrator,spp", [("cornell", 0, 64), ("cornell", 0, 512), ("bunny", 0, 32), ("conference", 1, 32), ("cornell", 2, 16), ("textured", 0, 32)])
def test_converged_image_relmse(name, integrator, spp, gpu_ctx):
    """relMSE = mean((a-b)^2 / (b^2 + 1e-3)) in linear RGB: the GPU image is as close to an
    oracle image as another oracle image with a different seed is (factor 1.5), and mean luminance agrees
    to 1 %.  The oracle arms use the reference schedule (per-tile xorshift, rng_mode=0)."""
    from lumo_b200 import native
    prog, blob, _ = small_scene(name)
    O = oracle_lib.OracleScene(prog)
    G = native.GpuScene(gpu_ctx, blob)
    def full(px, sp):      # Film::rgb_image before the transfer function (film.rs:173-193): direct + splats / (spp * filter integral)
        from lumo_b200 import PixelFilter
        return _rgb(px) + sp / (spp * PixelFilter.default().integral())
    ra = O.render(integrator=integrator, spp=spp, seed=11, rng_mode=0); rb = O.render(integrator=integrator, spp=spp, seed=12, rng_mode=0)
    rg = G.render(integrator=integrator, spp=spp, seed=13)
    a, b, g = full(ra[0], ra[1]), full(rb[0], rb[1]), full(rg[0], rg[1])
    ref = _relmse(a, b)
    assert _relmse(g, a) <= 1.5 * ref + 1e-6, (_relmse(g, a), ref)
    assert _relmse(g, b) <= 1.5 * ref + 1e-6, (_relmse(g, b), ref)
    m = 0.5 * (a.mean() + b.mean())
    tol = 0.08 if integrator == 2 else 0.01      # BDPT means are firefly-dominated at these sample counts: oracle seeds differ by +-5 %
    assert abs(g.mean() - m) <= tol * m + 2 * abs(a.mean() - b.mean()), (g.mean(), a.mean(), b.mean())
    G.close(); O.close()


def test_sample_ranges_compose(gpu_ctx):
    """G6 in one process: disjoint sample ranges sum to the full render (what the multi-GPU path relies on)."""
    from lumo_b200 import native
    prog, blob, _ = small_scene("cornell")
    G = native.GpuScene(gpu_ctx, blob)
    full, _, cf, _, _ = G.render(integrator=0, spp=8, seed=3)
    a, _, ca, _, _ = G.render(integrator=0, spp=8, seed=3, spp_begin=0, spp_end=3)
    b, _, cb, _, _ = G.render(integrator=0, spp=8, seed=3, spp_begin=3, spp_end=8)
    assert np.allclose(a + b, full, rtol=1e-12, atol=1e-300)
    assert ca["closest"] + cb["closest"] == cf["closest"] and ca["cost"] + cb["cost"] == cf["cost"]
    small, _, _, _, _ = G.render(integrator=0, spp=8, seed=3, wave_paths=4096)   # wave size must not matter
    assert np.allclose(small, full, rtol=1e-12, atol=1e-300)
    G.close()


def test_render_rejects_bad_arguments(gpu_ctx):
    from lumo_b200 import native
    prog, blob, _ = small_scene("cornell")
    G = native.GpuScene(gpu_ctx, blob)
    with pytest.raises(RuntimeError):
        G.render(integrator=7, spp=1)
    with pytest.raises(RuntimeError):
        G.render(integrator=0, spp=4, spp_begin=3, spp_end=2)
    with pytest.raises(RuntimeError):
        G.render(integrator=0, spp=1, rr_delta=-1.0)
    G.close()


VARIANTS = {
    "thin_lens": dict(cam=lambda b: b.lens_radius(8.0).focal_length(900.0)),
    "orthographic": dict(cam=lambda b: b.camera_type(1).zoom(0.004)),
    "mitchell_filter": dict(filt=("mitchell", 2.0, 1.0 / 3.0)),
    "triangle_filter": dict(filt=("triangle", 1.5, 0.0)),
    "srgb_space": dict(cam=lambda b: b.color_space(0)),
    "reinhard": dict(render=dict(tone_map=2)),
    "clamp": dict(render=dict(tone_map=1, tone_map_arg=0.5)),
    "uniform_sampler": dict(render=dict(sampler=0)),
    "jittered_sampler": dict(render=dict(sampler=1)),
    "sobol_sampler": dict(render=dict(sampler=3)),          # SamplerType::Sobol (samplers.rs:193-247)
}


@pytest.mark.parametrize("variant", list(VARIANTS))
def test_camera_film_sampler_variants(variant, gpu_ctx):
    """Camera models (camera.rs:221-268), pixel filters (filter.rs:82-102), colour space, per-sample tone maps
    (tone_mapping.rs:38-63) and samplers (samplers.rs:54-134) other than the defaults, against the oracle on shared streams."""
    from lumo_b200 import native, Scene, CameraBuilder, PixelFilter, illuminants
    v = VARIANTS[variant]
    b = (CameraBuilder.new().origin(278.0, 273.0, -800.0).towards(278.0, 273.0, 0.0).zoom(2.8).focal_length(0.035)
         .resolution((48, 48)).illuminant(illuminants.CORNELL))
    if "cam" in v: b = v["cam"](b)
    if "filt" in v:
        k, r, p = v["filt"]
        b = b.pixel_filter(getattr(PixelFilter, k)(r, p) if k == "mitchell" else getattr(PixelFilter, k)(r))
    cam = b.build()
    prog = Scene.cornell_box()._program(cam); blob = native.build_blob(prog)
    O = oracle_lib.OracleScene(prog); G = native.GpuScene(gpu_ctx, blob)
    kw = dict(integrator=0, spp=4, seed=21, rr_delta=0.05); kw.update(v.get("render", {}))
    epx, esp, ecnt, _ = O.render(rng_mode=1, **kw)
    gpx, gsp, gcnt, _, _ = G.render(**kw)
    assert gcnt["camera_paths"] == ecnt["camera_paths"] and gcnt["nonfinite"] == 0
    assert np.allclose(gpx[..., 3], epx[..., 3], rtol=1e-11, atol=1e-14), variant          # filter weights: same sample positions, same filter
    assert gcnt["closest"] == ecnt["closest"] and gcnt["cost"] == ecnt["cost"], (variant, gcnt, ecnt)
    scale = np.abs(epx[..., :3]).max()
    bad = (np.abs(gpx[..., :3] - epx[..., :3]) > 1e-11 * (np.abs(epx[..., :3]) + 1e-3 * scale)).any(axis=-1)   # filtered sums of both signs (negative filter lobes, XYZ -> RGB)
    assert not bad.any(), (variant, int(bad.sum()))
    G.close(); O.close()


def test_bdpt_batching_is_invisible(gpu_ctx):
    """BDPT runs in batches of samples (vertex buffers in HBM, up to 2^20 samples; LUMO_BDPT_BATCH overrides the cap).  The batch
    size must not change anything but the order of the film's atomic adds: same counters, same image to rounding."""
    import os
    from lumo_b200 import native
    prog, blob, _ = small_scene("caustics")
    G = native.GpuScene(gpu_ctx, blob)
    ref = G.render(integrator=2, spp=3, seed=9)
    for cap in ("1024", "5000"):
        os.environ["LUMO_BDPT_BATCH"] = cap
        try:
            got = G.render(integrator=2, spp=3, seed=9)
        finally:
            del os.environ["LUMO_BDPT_BATCH"]
        for k in ("camera_paths", "closest", "occlusion", "cost", "max_depth", "nonfinite"):
            assert got[2][k] == ref[2][k], (cap, k, got[2][k], ref[2][k])
        assert got[2]["iterations"] > ref[2]["iterations"]                      # it really ran in more batches
        assert np.allclose(got[0], ref[0], rtol=1e-9, atol=1e-12) and np.allclose(got[1], ref[1], rtol=1e-9, atol=1e-12), cap
    G.close()
