"""The reference-facing API end to end: `Renderer::new(scene, camera).samples(..).integrator(..).seed(..).sampler(..).tone_map(..)
.render()` -> `Film` -> `Film::save` (src/renderer.rs:24-99,159-244; src/tracer/film.rs:173-210), as a user of lumo writes it
(examples/cornell.rs).  CPU: the builder surface and its defaults.  GPU: the film of `render()` against the oracle's render of
the same scene program with the same streams, the saved PNG, and the error path (a failing render must not take the process
down when its context goes away)."""
import numpy as np
import pytest
import oracle_lib
import lumo_b200
from lumo_b200 import Renderer, Scene, Camera, Integrator, SamplerType, ToneMap


def _cornell(res=(48, 48)):
    cam = Camera.cornell_box(); cam._resolution = res
    return Scene.cornell_box(), cam


def test_builder_surface_and_defaults():
    scene, cam = _cornell()
    r = Renderer.new(scene, cam)
    assert r.num_samples == 1 and r._integrator == Integrator.PathTrace and r._sampler == SamplerType.MultiJittered      # renderer.rs:20-21,50; SURVEY A.17
    assert r._tone_map.kind == ToneMap.NoMap.kind and r._threads == 4
    r2 = r.samples(8).integrator(Integrator.DirectLight).seed(5).sampler(SamplerType.Sobol).threads(2).tone_map(ToneMap.Reinhard)
    assert r2 is r and r.num_samples == 8 and r._seed == 5 and r._sampler == SamplerType.Sobol == 3
    empty = Scene()
    with pytest.raises(AssertionError):
        Renderer.new(empty, cam)                                  # renderer.rs:42: a scene needs a light


@pytest.mark.gpu
@pytest.mark.parametrize("integrator,sampler", [(Integrator.PathTrace, SamplerType.MultiJittered), (Integrator.DirectLight, SamplerType.Sobol), (Integrator.BDPathTrace, SamplerType.Jittered)])
def test_render_returns_the_oracles_film(integrator, sampler, tmp_path):
    scene, cam = _cornell()
    r = Renderer.new(scene, cam).samples(4).integrator(integrator).seed(11).sampler(sampler)
    r.quiet = True
    film = r.render()
    O = oracle_lib.OracleScene(scene._program(cam))
    epx, esp, ecnt, _ = O.render(integrator=integrator, spp=4, seed=11, sampler=sampler, rng_mode=1)
    O.close()
    assert film.counters["camera_paths"] == ecnt["camera_paths"] == 4 * 48 * 48
    assert film.counters["closest"] == ecnt["closest"] and film.counters["cost"] == ecnt["cost"]
    scale = np.abs(epx).max()
    assert np.allclose(film.pixels, epx, rtol=1e-9, atol=1e-12 * scale) and np.allclose(film.splats, esp, rtol=1e-9, atol=1e-12 * max(np.abs(esp).max(), 1e-300))
    # Film::save: an 8-bit RGB PNG of Film::rgb_image
    path = tmp_path / "cornell.png"
    film.save(str(path))
    raw = path.read_bytes()
    assert raw[:8] == b"\x89PNG\r\n\x1a\n" and int.from_bytes(raw[16:20], "big") == 48 and int.from_bytes(raw[20:24], "big") == 48
    img = film.rgb_image()
    assert img.shape == (48, 48, 3) and img.dtype == np.uint8 and img.max() > 0
    from lumo_b200 import image
    back = image.decode_png(raw) if hasattr(image, "decode_png") else None
    if back is not None:
        assert np.array_equal(np.asarray(back)[..., :3].reshape(48, 48, 3), img)


@pytest.mark.gpu
def test_failed_render_leaves_the_process_usable():
    """A render that the library refuses (Sobol beyond its 1023 points) raises; scene and context are released in order."""
    scene, cam = _cornell((16, 16))
    r = Renderer.new(scene, cam).samples(2000).sampler(SamplerType.Sobol).seed(1)
    r.quiet = True
    with pytest.raises(RuntimeError, match="Sobol"):
        r.render()
    ok = Renderer.new(scene, cam).samples(2).seed(1); ok.quiet = True
    assert ok.render().counters["camera_paths"] == 2 * 16 * 16
