"""Bidirectional subpaths beyond the in-place vertex block: the reference grows a `Vec<Vertex>` up to BDPT_MAX_DEPTH = 1024
(bd_path_trace.rs:7, path_gen.rs:139); the device keeps 64 vertices per subpath in place and takes further blocks of 64
from an overflow pool (bdpt.cuh), and the MIS weight (mis.rs:103-239) runs over the vertices directly instead of over
fixed-size arrays.  A hall of mirrors with a tiny Russian-roulette threshold gives subpaths of a few hundred vertices."""
import numpy as np
import pytest
import oracle_lib
from lumo_b200 import native, Scene, Material, Rectangle, CameraBuilder, Spectrum

pytestmark = pytest.mark.gpu


def _hall_of_mirrors(res=(10, 8)):
    LIGHT_EPS = 0.001
    ground = -0.8; ceiling = -ground; right = 1.0; left = -right; front = -2.0; back = 0.0; l_dim = 0.25
    s = Scene()
    s.add_light(Rectangle((-l_dim, ceiling - LIGHT_EPS, 0.6 * front + l_dim), (-l_dim, ceiling - LIGHT_EPS, 0.6 * front - l_dim),
                          (l_dim, ceiling - LIGHT_EPS, 0.6 * front - l_dim), Material.light(Spectrum.from_srgb(252, 201, 138))))
    mirror = Material.mirror()
    s.add(Rectangle((left, ground, back), (left, ground, front), (left, ceiling, front), mirror))
    s.add(Rectangle((right, ground, front), (right, ground, back), (right, ceiling, back), mirror))
    s.add(Rectangle((left, ground, back), (right, ground, back), (right, ground, front), Material.diffuse(Spectrum.from_srgb(200, 200, 200))))
    s.add(Rectangle((left, ceiling, front), (right, ceiling, front), (right, ceiling, back), mirror))
    s.add(Rectangle((left, ground, front), (right, ground, front), (right, ceiling, front), mirror))
    s.add(Rectangle((left, ground, back), (left, ceiling, back), (right, ceiling, back), mirror))          # closes the box behind the camera
    cam = CameraBuilder.new().origin(0.0, 0.0, -0.2).towards(0.3, 0.1, -2.0).resolution(res).build()
    return s, cam


def test_long_subpaths_match_the_oracle_and_nothing_is_cut(gpu_ctx):
    scene, cam = _hall_of_mirrors()
    prog = scene._program(cam); blob = native.build_blob(prog)
    O = oracle_lib.OracleScene(prog); G = native.GpuScene(gpu_ctx, blob)
    kw = dict(integrator=2, spp=2, seed=19, rr_delta=1e-9)
    epx, esp, ecnt, _ = O.render(rng_mode=1, **kw)
    gpx, gsp, gcnt, _, _ = G.render(**kw)
    assert gcnt["max_depth"] > 64, gcnt                                   # beyond the in-place block: the overflow pool is in use
    assert gcnt["bdpt_cut_subpaths"] == 0 and gcnt["nonfinite"] == 0
    for k in ("camera_paths", "closest", "occlusion", "cost"):
        assert gcnt[k] == ecnt[k], (k, gcnt[k], ecnt[k])
    scale = max(np.abs(epx).max(), 1e-300)
    assert np.allclose(gpx, epx, rtol=1e-9, atol=1e-12 * scale)
    assert np.allclose(gsp, esp, rtol=1e-9, atol=1e-12 * max(np.abs(esp).max(), 1e-300))
    # PathTrace through the same hall: depth is unbounded there too
    e2 = O.render(rng_mode=1, integrator=0, spp=2, seed=19, rr_delta=1e-9); g2 = G.render(integrator=0, spp=2, seed=19, rr_delta=1e-9)
    assert g2[2]["max_depth"] > 64 and g2[2]["closest"] == e2[2]["closest"]
    assert np.allclose(g2[0], e2[0], rtol=1e-9, atol=1e-12 * max(np.abs(e2[0]).max(), 1e-300))
    G.close(); O.close()
