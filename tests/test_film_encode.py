"""Film finalisation on the device (lumo_gpu_film_encode / _dev) against a CPU restatement of
Film::rgb_image (src/tracer/film.rs:173-193), Pixel::value (film.rs:82-90) and TransferFunction::apply
(src/tracer/color/space.rs:8-36).  Every byte must be equal: the transfer curve's pow is csrc/common/lumo_math.h on
both sides (the oracle's build of it here), so there is no pow-rounding slack left (round 1 allowed one code on 2e-4
of the bytes, CUDA pow vs glibc pow)."""
import math
import numpy as np
import pytest
from conftest import small_scene


def _pow(c, y):
    """lm_pow as compiled for the oracle (oracle_math_eval fn 8)."""
    import ctypes as C
    import oracle_lib
    x = np.array([c], np.float64); yy = np.array([y], np.float64); out = np.zeros(1)
    oracle_lib.lib().oracle_math_eval(C.c_int(8), x.ctypes.data_as(C.POINTER(C.c_double)), yy.ctypes.data_as(C.POINTER(C.c_double)), C.c_uint64(1), out.ctypes.data_as(C.POINTER(C.c_double)))
    return float(out[0])


def _apply_scalar(c, transfer):
    """TransferFunction::apply with Rust's saturating float -> u8 cast (NaN -> 0)."""
    if transfer == 1:
        beta = 0.018053968510807; alpha = 1.0 + 5.5 * beta
        ec = 4.5 * c if c <= beta else (alpha * _pow(c, 0.45) - (alpha - 1.0) if not math.isnan(c) else c)
    else:
        ec = 12.92 * c if c <= 0.0031308 else (1.055 * _pow(c, 1.0 / 2.4) - 0.055 if not math.isnan(c) else c)
    v = ec * 255.0
    if math.isnan(v) or v <= 0.0: return 0
    return 255 if v >= 255.0 else int(v)


def rgb_image_ref(pixels, splats, splat_scale, integral, transfer):
    """Film::rgb_image, pixel by pixel (film.rs:176-189)."""
    px = pixels.reshape(-1, 4); sp = splats.reshape(-1, 3)
    out = np.zeros((len(px), 3), dtype=np.uint8)
    with np.errstate(invalid="ignore", divide="ignore"):
        lin = px[:, :3] / px[:, 3:4] + splat_scale * sp / integral
    for i in range(len(px)):
        for k in range(3):
            out[i, k] = _apply_scalar(float(lin[i, k]), transfer)
    return out.reshape(splats.shape)


def _accumulators(h, w, seed):
    rs = np.random.RandomState(seed)
    wsum = rs.rand(h, w, 1) * 40.0 + 0.5
    val = np.exp(rs.randn(h, w, 3) * 2.0 - 2.0)             # linear values from ~1e-4 to ~10: both curve segments + saturation
    px = np.concatenate([val * wsum, wsum], -1)
    sp = np.where(rs.rand(h, w, 3) < 0.3, rs.rand(h, w, 3) * 5.0, 0.0)
    flat = px.reshape(-1, 4)
    flat[0] = 0.0                                           # pixel that never received a sample: 0/0 -> NaN -> 0
    flat[1, :3] = -flat[1, :3]                              # negative radiance -> 0
    flat[2, :3] = 1e300; flat[3, 0] = np.inf; flat[4, 1] = np.nan
    flat[5] = (0.0031308 * 2.0, 0.0031307 * 2.0, 0.0, 2.0)  # either side of the sRGB knee
    sp.reshape(-1, 3)[:6] = 0.0
    return px, sp


def _compare(got, want, max_off_fraction=0.0):
    assert got.shape == want.shape and got.dtype == np.uint8
    assert np.array_equal(got, want), "bytes differ: %d of %d" % (int((got != want).sum()), got.size)


def test_reference_restatement_matches_host_film():
    """CPU: the scalar restatement above and the vectorised host finalisation (lumo_b200/film.py) agree byte for byte."""
    from lumo_b200 import color
    px, sp = _accumulators(37, 53, 3)
    for transfer, cs in ((0, 1), (1, 2)):
        want = rgb_image_ref(px, sp, 1.0 / 16, 1.7, transfer)
        with np.errstate(invalid="ignore", divide="ignore"):
            got = color.encode(px[..., :3] / px[..., 3:4] + (1.0 / 16) * sp / 1.7, cs)
        assert np.array_equal(got, want)
        assert tuple(want.reshape(-1, 3)[0]) == (0, 0, 0) and tuple(want.reshape(-1, 3)[1]) == (0, 0, 0)
        assert tuple(want.reshape(-1, 3)[2]) == (255, 255, 255)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(1, 1), (1, 3), (5, 7), (37, 53), (96, 128)])
@pytest.mark.parametrize("transfer", [0, 1])
def test_film_encode_matches_restatement(shape, transfer, gpu_ctx):
    px, sp = _accumulators(max(shape[0], 2), max(shape[1], 3), 11 + shape[0])
    px, sp = np.ascontiguousarray(px[:shape[0], :shape[1]]), np.ascontiguousarray(sp[:shape[0], :shape[1]])
    got = gpu_ctx.film_encode(px, sp, 1.0 / 64, 2.25, transfer)
    _compare(got, rgb_image_ref(px, sp, 1.0 / 64, 2.25, transfer), max_off_fraction=1e-3 if px.size < 4000 else 2e-4)
    # linear segment and the special values are exact
    n = 6 if shape[0] >= 2 and shape[1] >= 3 else 1
    assert np.array_equal(got.reshape(-1, 3)[:n], rgb_image_ref(px, sp, 1.0 / 64, 2.25, transfer).reshape(-1, 3)[:n])


@pytest.mark.gpu
def test_film_encode_full_hd_and_rendered_film(gpu_ctx):
    """1920x1080 accumulators against the vectorised host finalisation; then a rendered cornell film through
    Film.rgb_image_device vs Film.rgb_image."""
    from lumo_b200 import color, native
    from lumo_b200.film import Film
    from lumo_b200 import PixelFilter
    px, sp = _accumulators(1080, 1920, 5)
    with np.errstate(invalid="ignore", divide="ignore"):
        want = color.encode(px[..., :3] / px[..., 3:4] + (1.0 / 1024) * sp / 1.3, 1)
    _compare(gpu_ctx.film_encode(px, sp, 1.0 / 1024, 1.3, 0), want, max_off_fraction=1e-5)
    prog, blob, ig = small_scene("cornell")
    G = native.GpuScene(gpu_ctx, blob)
    gpx, gsp, cnt, _, _ = G.render(integrator=2, spp=4, seed=3)        # BDPT: splats are populated
    G.close()
    film = Film(gpx, gsp, 4, PixelFilter.default(), 1)
    dev = film.rgb_image_device(ctx=gpu_ctx)
    _compare(dev, film.rgb_image(), max_off_fraction=1e-3)
    assert dev.max() > 0


@pytest.mark.gpu
def test_film_encode_dev_reads_render_dev_buffers(gpu_ctx):
    """render_dev leaves the accumulators in device memory; film_encode_dev finishes them there."""
    torch = pytest.importorskip("torch")
    from lumo_b200 import color, native
    prog, blob, ig = small_scene("bunny")
    G = native.GpuScene(gpu_ctx, blob)
    W, H = G.res_x, G.res_y
    dpx = torch.zeros(H, W, 4, dtype=torch.float64, device="cuda:0"); dsp = torch.zeros(H, W, 3, dtype=torch.float64, device="cuda:0")
    torch.cuda.synchronize()
    G.render_dev(dpx.data_ptr(), dsp.data_ptr(), integrator=0, spp=4, seed=9)
    rgb, ms = gpu_ctx.film_encode_dev(dpx.data_ptr(), dsp.data_ptr(), (H, W), 0.25, 1.0, 0)
    G.close()
    px, sp = dpx.cpu().numpy(), dsp.cpu().numpy()
    with np.errstate(invalid="ignore", divide="ignore"):
        want = color.encode(px[..., :3] / px[..., 3:4] + 0.25 * sp / 1.0, 1)
    _compare(rgb, want, max_off_fraction=1e-3)
    assert ms > 0.0 and rgb.max() > 0


@pytest.mark.gpu
def test_film_encode_rejects_bad_arguments(gpu_ctx):
    px, sp = _accumulators(4, 4, 1)
    with pytest.raises(RuntimeError):
        gpu_ctx.film_encode(px, sp, 1.0, 1.0, transfer=7)


@pytest.mark.gpu
@pytest.mark.parametrize("transfer", [0, 1])
def test_film_encode_exact_next_to_code_boundaries(transfer, gpu_ctx):
    """The kernel settles a code from an f32 estimate of the curve unless the value is next to a code boundary; this
    walks every boundary k/255 at offsets from 1e-2 down to 1e-9 (both sides) and the saturation edge, where the bytes
    must equal the f64 restatement exactly."""
    ks = np.arange(1, 256, dtype=np.float64)
    offs = np.array([0.0, 1e-2, 2e-3, 1.4e-3, 1e-3, 3e-4, 1e-4, 1e-5, 1e-6, 1e-7, 1e-9])
    v = (ks[:, None, None] + np.stack([offs, -offs], -1)[None]).reshape(-1) / 255.0          # target ec
    if transfer == 1:
        beta = 0.018053968510807; alpha = 1.0 + 5.5 * beta
        c = np.where(v <= 4.5 * beta, v / 4.5, ((v + (alpha - 1.0)) / alpha) ** (1.0 / 0.45))
    else:
        c = np.where(v <= 12.92 * 0.0031308, v / 12.92, ((v + 0.055) / 1.055) ** 2.4)
    c = np.concatenate([c, [1.0, 1.0 + 1e-12, 1.0 - 1e-12, 1.01, 0.999, 3.0, 1e30]])
    n = (len(c) + 2) // 3 * 3
    c = np.concatenate([c, np.full(n - len(c), 0.5)])
    px = np.concatenate([c.reshape(-1, 1, 3), np.ones((n // 3, 1, 1))], -1)                   # weight 1: value = c exactly
    sp = np.zeros((n // 3, 1, 3))
    got = gpu_ctx.film_encode(px, sp, 1.0, 1.0, transfer)
    want = rgb_image_ref(px, sp, 1.0, 1.0, transfer)
    assert np.array_equal(got, want), np.nonzero((got != want).reshape(-1))[0][:10]      # the same pow on both sides: equal even on the boundary itself


@pytest.mark.gpu
@pytest.mark.parametrize("transfer", [0, 1])
def test_f32_prepass_never_changes_a_byte(transfer, gpu_ctx, monkeypatch):
    """The kernel's f32 pre-pass against the same kernel with every pixel on the f64 path (LUMO_FILM_F64=1): identical
    bytes over 2 M random pixels (with splats, zero weights, negatives, NaN, inf) and over weights / values at the edges
    of the f32 range."""
    px, sp = _accumulators(1080, 1920, 21 + transfer)
    flat = px.reshape(-1, 4); rs = np.random.RandomState(4)
    flat[100:200, 3] = 1e-30; flat[200:300, 3] = 1e25; flat[300:400, :3] *= 1e-40; flat[400:500, :3] *= 1e45      # denormal / huge in f32
    flat[500:600] *= 1e-25; flat[600:700] *= 1e22
    flat[700:800, :3] = -np.abs(flat[700:800, :3]) * 1e-3; sp.reshape(-1, 3)[700:800] = rs.rand(100, 3) * 60.0         # negative direct + positive splat
    for scale, integral in ((1.0 / 64, 0.9977), (1.0, 1e-12), (1e12, 1.0)):
        monkeypatch.delenv("LUMO_FILM_F64", raising=False)
        fast = gpu_ctx.film_encode(px, sp, scale, integral, transfer)
        monkeypatch.setenv("LUMO_FILM_F64", "1")
        exact = gpu_ctx.film_encode(px, sp, scale, integral, transfer)
        assert np.array_equal(fast, exact), (scale, integral, int((fast != exact).sum()))
    monkeypatch.delenv("LUMO_FILM_F64", raising=False)
    assert len(np.unique(fast)) > 200
