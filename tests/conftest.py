import os
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


# Small versions of the BASELINE configs: the oracle must finish each in seconds.
SMALL = {
    "cornell": dict(resolution=(64, 64)),
    "bunny": dict(n_tris=6000, resolution=(64, 48)),
    "dragon": dict(n_tris=12288, resolution=(64, 48)),
    "caustics": dict(n_tris=2000, resolution=(64, 48)),
    "conference": dict(n_chunks=8, tris_per_chunk=256, resolution=(64, 48)),
    "bistro": dict(n_chunks=16, tris_per_chunk=256, n_emissive=32, resolution=(64, 36)),
    "textured": dict(n_tris=512, resolution=(64, 48)),
}

_cache = {}


def small_scene(name, box_filter=False):
    """(program bytes, blob bytes, integrator) of the small version of a config, built once.
    box_filter: PixelFilter::square(0.5), i.e. a pixel only receives its own samples."""
    key = (name, box_filter)
    if key not in _cache:
        from lumo_b200 import scenes, native, PixelFilter
        s, cam, ig = scenes.CONFIGS[name](**SMALL[name])
        if box_filter:
            cam._pixel_filter = PixelFilter.square(0.5)
        prog = s._program(cam)
        _cache[key] = (prog, native.build_blob(prog), ig)
    return _cache[key]


def ray_batches(oracle_scene, n, seed):
    """(i) primary rays from the config camera, (ii) incoherent rays inside the scene bounds,
    (iii) secondary rays: bounce (i) off the first hit with a random direction."""
    rs = np.random.RandomState(seed)
    W, H = oracle_scene.res_x, oracle_scene.res_y
    raster = rs.rand(n, 2) * np.array([W, H])
    o1, d1 = oracle_scene.camera_rays(raster, rs.rand(n, 2))
    b = oracle_scene.bounds()
    lo, hi = b[:3], b[3:]
    o2 = lo + rs.rand(n, 3) * (hi - lo)
    z = 1.0 - 2.0 * rs.rand(n); ph = 2 * np.pi * rs.rand(n); rr = np.sqrt(np.maximum(1 - z * z, 0))
    d2 = np.stack([rr * np.cos(ph), rr * np.sin(ph), z], -1)
    _, _, t, _ = oracle_scene.trace_closest(o1, d1)
    ok = np.isfinite(t)
    o3 = o1[ok] + d1[ok] * (t[ok] * (1 - 1e-9))[:, None]
    z = 1.0 - 2.0 * rs.rand(len(o3)); ph = 2 * np.pi * rs.rand(len(o3)); rr = np.sqrt(np.maximum(1 - z * z, 0))
    d3 = np.stack([rr * np.cos(ph), rr * np.sin(ph), z], -1)
    return {"primary": (o1, d1), "incoherent": (o2, d2), "secondary": (o3, d3)}


@pytest.fixture(scope="session")
def gpu_ctx():
    from lumo_b200 import native
    ctx = native.GpuContext(0)
    yield ctx
    ctx.close()
