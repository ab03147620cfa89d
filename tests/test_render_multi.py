"""lumo_gpu_render_multi (one process, n GPUs): the sample range is split over the scenes' contexts and the films are
summed by the peer-memory reduce kernel.  Because every path's random stream is keyed by (pixel, global sample index),
the result must equal the single-context render up to the summation order of the film atomics.  On a
one-GPU box the contexts all live on device 0 (the same code path: threads, range split, reduce kernel); with two or
more GPUs the second context sits on device 1 and the reduce reads its film over NVLink."""
import numpy as np
import pytest
from conftest import small_scene


def _device_count():
    import ctypes as C
    from lumo_b200 import native
    n = C.c_int32(0)
    native._check(native.gpu_lib().lumo_gpu_device_count(C.byref(n)), "lumo_gpu_device_count")
    return n.value


@pytest.mark.gpu
@pytest.mark.parametrize("name,integrator,spp,n_ctx", [("cornell", 0, 8, 2), ("cornell", 0, 7, 3), ("bunny", 1, 4, 2), ("cornell", 2, 5, 2), ("cornell", 0, 1, 2)])
def test_render_multi_equals_single_render(name, integrator, spp, n_ctx, gpu_ctx):
    from lumo_b200 import native
    prog, blob, _ = small_scene(name)
    G = native.GpuScene(gpu_ctx, blob)
    full_px, full_sp, cf, _, _ = G.render(integrator=integrator, spp=spp, seed=3)
    ndev = _device_count()
    ctxs = [native.GpuContext(g % ndev) for g in range(n_ctx)]
    scenes = [native.GpuScene(c, blob) for c in ctxs]
    try:
        px, sp, cm, deltas, ms = native.render_multi(scenes, integrator=integrator, spp=spp, seed=3)
    finally:
        for s in scenes: s.close()
        for c in ctxs: c.close()
    G.close()
    assert np.allclose(px, full_px, rtol=1e-12, atol=1e-300) and np.allclose(sp, full_sp, rtol=1e-12, atol=1e-300)
    for k in ("camera_paths", "closest", "occlusion", "cost"):
        assert cm[k] == cf[k], k
    assert cm["camera_paths"] == spp * G.res_x * G.res_y and ms > 0.0 and px[..., 3].min() > 0.0


@pytest.mark.gpu
def test_render_multi_on_two_gpus_if_present(gpu_ctx):
    if _device_count() < 2:
        pytest.skip("one GPU on this box")
    from lumo_b200 import native
    prog, blob, _ = small_scene("bunny")
    G = native.GpuScene(gpu_ctx, blob)
    full_px, full_sp, cf, _, _ = G.render(integrator=0, spp=8, seed=5)
    G.close()
    ctxs = [native.GpuContext(0), native.GpuContext(1)]
    scenes = [native.GpuScene(c, blob) for c in ctxs]
    px, sp, cm, _, _ = native.render_multi(scenes, integrator=0, spp=8, seed=5)
    for s in scenes: s.close()
    for c in ctxs: c.close()
    assert np.allclose(px, full_px, rtol=1e-12, atol=1e-300) and cm["closest"] == cf["closest"]


@pytest.mark.gpu
def test_render_multi_rejects_bad_arguments(gpu_ctx):
    from lumo_b200 import native
    prog, blob, _ = small_scene("cornell")
    a = native.GpuScene(gpu_ctx, blob); b = native.GpuScene(gpu_ctx, blob)
    with pytest.raises(RuntimeError, match="own context"):
        native.render_multi([a, b], spp=2)
    with pytest.raises(RuntimeError, match="null pointer"):
        native._check(native.gpu_lib().lumo_gpu_render_multi(None, 0, None, None), "lumo_gpu_render_multi")
    c2 = native.GpuContext(0)
    other = native.GpuScene(c2, small_scene("bunny")[1])
    with pytest.raises(RuntimeError, match="resolution"):
        native.render_multi([a, other], spp=2)
    twin = native.GpuScene(c2, blob)
    with pytest.raises(RuntimeError, match="unknown integrator"):
        native.render_multi([a, twin], integrator=9, spp=2)
    px, _, cnt, _, _ = native.render_multi([a], spp=2)                 # n = 1 is the plain render
    assert cnt["camera_paths"] == 2 * a.res_x * a.res_y and px[..., 3].min() > 0.0
    twin.close(); other.close(); c2.close(); a.close(); b.close()
