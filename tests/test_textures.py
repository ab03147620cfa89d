"""Textures and bump maps (src/tracer/texture.rs:23-113, src/perlin.rs, src/image.rs:99-193,
material.rs:324-331).  The reference has no test or golden vector for them ("parity unpinned"): the oracle restatement is
checked here against the properties the formulas imply and against independent numpy restatements; the GPU is checked
against the oracle through the render-parity tests on `scenes.textured` (tests/test_render_parity.py) and below."""
import numpy as np
import pytest
import oracle_lib
from conftest import small_scene
from lumo_b200 import Scene, Material, Texture, Spectrum, Rectangle, Image, native
from lumo_b200 import program as P
from lumo_b200.api import CameraBuilder
from lumo_b200.image import decode_png, encode_png
import struct, zlib


def _scene_with(textures):
    """a one-wall scene whose program carries `textures` (each on its own diffuse rectangle)"""
    s = Scene()
    s.add_light(Rectangle((-1, 1, -1), (-1, 1, -2), (1, 1, -2), Material.light(Spectrum.WHITE())))
    for i, t in enumerate(textures):
        s.add(Rectangle((-1, -1 + 0.1 * i, 0), (1, -1 + 0.1 * i, 0), (1, -1 + 0.1 * i, -2), Material.diffuse(t)))
    return s._program(CameraBuilder.new().resolution((16, 16)).build())


def _tex_ids(O, kind):
    return [i for i in range(O.texture_count()) if O.texture_kind(i) == kind]


def test_png_decode_rgb_grey_palette():
    rs = np.random.RandomState(1)
    rgb = rs.randint(0, 256, size=(5, 7, 3)).astype(np.uint8)
    assert np.array_equal(decode_png(encode_png(rgb)), rgb)

    def png(w, h, depth, ctype, rows, plte=None, filt=0):
        def chunk(tag, body): return struct.pack(">I", len(body)) + tag + body + struct.pack(">I", zlib.crc32(tag + body) & 0xFFFFFFFF)
        raw = b"".join(bytes([filt]) + bytes(r) for r in rows)
        out = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, 0))
        if plte is not None: out += chunk(b"PLTE", bytes(plte))
        return out + chunk(b"IDAT", zlib.compress(raw)) + chunk(b"IEND", b"")
    # grey + alpha: alpha dropped, grey replicated (image.rs:52-59)
    g = decode_png(png(2, 1, 8, 4, [[10, 255, 200, 0]]))
    assert g.tolist() == [[[10, 10, 10], [200, 200, 200]]]
    # RGBA (image.rs:60-72)
    assert decode_png(png(1, 1, 8, 6, [[1, 2, 3, 4]])).tolist() == [[[1, 2, 3]]]
    # 8-bit palette
    pal = [255, 0, 0, 0, 255, 0, 0, 0, 255]
    assert decode_png(png(3, 1, 8, 3, [[2, 0, 1]], plte=pal)).tolist() == [[[0, 0, 255], [255, 0, 0], [0, 255, 0]]]
    # 4-bit palette: the reference takes the LOW nibble first (image.rs:39-41)
    assert decode_png(png(2, 1, 4, 3, [[0x21]], plte=pal)).tolist() == [[[0, 255, 0], [0, 0, 255]]]
    # Sub / Up / Paeth filters undo correctly
    a = rs.randint(0, 256, size=(3, 4, 3)).astype(np.uint8)
    rows, prev = [], np.zeros(12, np.int32)
    for y, ft in enumerate((1, 2, 4)):
        cur = a[y].reshape(-1).astype(np.int32); enc = np.zeros(12, np.int32)
        for i in range(12):
            left = cur[i - 3] if i >= 3 else 0; up = prev[i]; ul = prev[i - 3] if i >= 3 else 0
            if ft == 1: pred = left
            elif ft == 2: pred = up
            else:
                pa, pb, pc = abs(up - ul), abs(left - ul), abs(left + up - 2 * ul)
                pred = left if (pa <= pb and pa <= pc) else (up if pb <= pc else ul)
            enc[i] = (cur[i] - pred) & 255
        rows.append((ft, enc.tolist())); prev = cur
    def chunk(tag, body): return struct.pack(">I", len(body)) + tag + body + struct.pack(">I", zlib.crc32(tag + body) & 0xFFFFFFFF)
    raw = b"".join(bytes([ft]) + bytes(r) for ft, r in rows)
    data = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", 4, 3, 8, 2, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(raw)) + chunk(b"IEND", b"")
    assert np.array_equal(decode_png(data), a)


def test_image_spectrum_and_bump_construction():
    rgb = np.array([[[255, 0, 0], [0, 255, 0]], [[0, 0, 255], [255, 255, 255]]], np.uint8)
    img = Image.from_rgb8(rgb)
    assert img.data.shape == (2, 2, 4) and img.width == 2 and img.height == 2
    # each pixel is Spectrum::from_srgb of its colour (image.rs:263-265); the mean is from_rgb of the mean linear colour
    for (y, x) in ((0, 0), (0, 1), (1, 0), (1, 1)):
        assert np.array_equal(img.data[y, x], np.array(Spectrum.from_srgb(*rgb[y, x]).as_tuple()))
    assert np.allclose(img.mean.as_tuple(), Spectrum.from_rgb(0.5, 0.5, 0.5).as_tuple())
    b = Image.bump_from_rgb8(np.array([[[128, 128, 255], [255, 128, 128]]], np.uint8))       # image.rs:158-170
    assert np.allclose(np.linalg.norm(b.data, axis=-1), 1.0, atol=1e-15)
    assert np.allclose(b.data[0, 0], [0, 0, 1]) and np.allclose(b.data[0, 1], np.array([127 / 128, 0, 0]) / (127 / 128))
    assert np.allclose(Image.mean_vec3_from_rgb8(np.full((3, 3, 3), 128, np.uint8)), 0.5)     # image.rs:82-97


def test_oracle_solid_checker_mandelbrot():
    white, red = Spectrum.from_srgb(242, 242, 242), Spectrum.RED()
    prog = _scene_with([Texture.Checkerboard(white, red, 4.0), Texture.Mandelbrot(),
                        Texture.Checkerboard(Texture.Checkerboard(white, red, 2.0), Spectrum.GREEN(), 8.0)])
    O = oracle_lib.OracleScene(prog)
    rs = np.random.RandomState(3)
    uv = rs.rand(2000, 2)
    lam_u = 0.37
    sol = {k: O.texture_eval(_tex_ids(O, P.TEX_SOLID)[k], uv[:1], lam_u)[0] for k in range(len(_tex_ids(O, P.TEX_SOLID)))}
    chk = _tex_ids(O, P.TEX_CHECKER)
    got = O.texture_eval(chk[0], uv, lam_u)
    even = (np.floor(uv[:, 0] * 4.0) + np.floor(uv[:, 1] * 4.0)).astype(np.int64) % 2 == 0              # texture.rs:66-73
    assert np.array_equal(got[even], np.broadcast_to(sol[0], got[even].shape)) and np.array_equal(got[~even], np.broadcast_to(sol[1], got[~even].shape))
    # nested: the children see the ORIGINAL uv
    inner_even = (np.floor(uv[:, 0] * 2.0) + np.floor(uv[:, 1] * 2.0)).astype(np.int64) % 2 == 0
    outer_even = (np.floor(uv[:, 0] * 8.0) + np.floor(uv[:, 1] * 8.0)).astype(np.int64) % 2 == 0
    got = O.texture_eval(chk[2], uv, lam_u)
    inner = O.texture_eval(chk[1], uv, lam_u)
    assert np.array_equal(got[outer_even], inner[outer_even])
    assert np.array_equal(inner[inner_even], O.texture_eval(chk[1], uv[inner_even], lam_u))
    # Mandelbrot (texture.rs:75-91): c = 2 (u - 0.75, v - 0.5); white inside the set
    mb = _tex_ids(O, P.TEX_MANDELBROT)[0]
    pts = np.array([[0.75, 0.5], [0.25, 0.5], [0.0, 0.0], [1.0, 1.0], [0.7, 0.55]])       # c = 0, -1, (-1.5,-1), (0.5,1), (-0.1,0.1)
    assert O.texture_eval(mb, pts, lam_u)[:, 0].tolist() == [1.0, 1.0, 0.0, 0.0, 1.0]
    c = 2.0 * (uv - np.array([0.75, 0.5])); z = np.zeros(len(uv), complex); cc = c[:, 0] + 1j * c[:, 1]; alive = np.ones(len(uv), bool)
    for _ in range(256):
        z = np.where(alive, z * z + cc, z); alive &= (z.real ** 2 + z.imag ** 2) < 4096.0
    agree = (O.texture_eval(mb, uv, lam_u)[:, 0] == 1.0) == alive
    assert agree.mean() > 0.995                                                            # boundary points may differ in the last bit
    O.close()


def test_oracle_perlin_and_marble():
    L = oracle_lib.lib()
    import ctypes as C
    tab = np.zeros(1536)
    L.oracle_perlin_tables(C.c_uint64(7), tab.ctypes.data_as(C.POINTER(C.c_double)))
    lattice, perms = tab[:768].reshape(256, 3), tab[768:].reshape(3, 256)
    assert np.allclose(np.linalg.norm(lattice, axis=1), 1.0, atol=1e-12)                   # square_to_sphere
    for p in perms: assert sorted(p.astype(int).tolist()) == list(range(256))             # gen_perm is a permutation
    # the xorshift stream: lattice[0] comes from the first two floats (rng.rs:71-75, maps.rs:49-55)
    xs = np.zeros(2, np.uint64); L.oracle_xorshift(C.c_uint64(7), C.c_uint64(2), xs.ctypes.data_as(C.POINTER(C.c_uint64)))
    u = np.minimum(xs.astype(np.float64) * 2.0 ** -64, 1.0 - 2.0 ** -52)
    z = 1.0 - 2.0 * u[1]; r = np.sqrt(max(1.0 - z * z, 0.0))
    assert np.allclose(lattice[0], [r * np.cos(2 * np.pi * u[0]), r * np.sin(2 * np.pi * u[0]), z], atol=1e-15)
    # noise: independent numpy restatement of perlin.rs:50-109
    rs = np.random.RandomState(5)
    pts = rs.rand(500, 3) * 40.0
    got = np.zeros(500); L.oracle_perlin_noise(C.c_uint64(7), pts.ctypes.data_as(C.POINTER(C.c_double)), C.c_uint64(500), got.ctypes.data_as(C.POINTER(C.c_double)))
    fl = np.floor(pts).astype(np.int64); w0 = pts - np.trunc(pts)
    w = ((6.0 * w0 - 15.0) * w0 + 10.0) * w0 * w0 * w0
    ref = np.zeros(500)
    for i in range(2):
        for j in range(2):
            for k in range(2):
                h = perms[0][(fl[:, 0] + i) % 256].astype(int) ^ perms[1][(fl[:, 1] + j) % 256].astype(int) ^ perms[2][(fl[:, 2] + k) % 256].astype(int)
                n = lattice[h]; idx = np.array([i, j, k], float)
                widx = 2.0 * w * idx + 1.0 - w - idx
                ref += widx[:, 0] * widx[:, 1] * widx[:, 2] * np.sum(n * (w - idx), axis=1)
    assert np.allclose(got, ref, rtol=1e-12, atol=1e-14)
    assert np.abs(got).max() < 1.8 and np.abs(got).mean() > 0.01
    # lattice points: every weight term vanishes -> noise 0 (w = 0: widx selects corner 0, dot(n, 0) = 0)
    lat = np.floor(pts[:50]); g0 = np.zeros(50)
    L.oracle_perlin_noise(C.c_uint64(7), lat.ctypes.data_as(C.POINTER(C.c_double)), C.c_uint64(50), g0.ctypes.data_as(C.POINTER(C.c_double)))
    assert np.all(g0 == 0.0)
    # Marble (texture.rs:56-64): colour * (1 - (0.5 + 0.5 sin(..))^6) in [0, 1] of the solid colour, deterministic in the seed
    prog = _scene_with([Texture.Marble(7, Spectrum.YELLOW()), Texture.Marble(8, Spectrum.YELLOW()),
                        Texture.Checkerboard(Spectrum.YELLOW(), Spectrum.YELLOW(), 1.0)])   # (Solid textures travel inline; a checkerboard child is a record)
    O = oracle_lib.OracleScene(prog)
    uv = rs.rand(1000, 2)
    m7, m8 = (O.texture_eval(t, uv) for t in _tex_ids(O, P.TEX_MARBLE))
    sol = O.texture_eval(_tex_ids(O, P.TEX_SOLID)[-1], uv[:1])[0]
    ratio = m7 / sol
    assert ratio.min() >= 0.0 and ratio.max() <= 1.0 and np.allclose(ratio, ratio[:, :1])   # one scalar per point
    assert 0.2 < ratio.mean() < 0.98 and not np.allclose(m7, m8)
    assert np.array_equal(m7, O.texture_eval(_tex_ids(O, P.TEX_MARBLE)[0], uv))
    O.close()


def test_oracle_image_bilinear_and_bump():
    rs = np.random.RandomState(9)
    pal = np.array([[230, 60, 50], [40, 170, 90], [60, 90, 220], [250, 250, 250]], np.uint8)
    rgb = pal[rs.randint(0, 4, size=(3, 5))]
    img = Image.from_rgb8(rgb)
    bump = Image.bump_from_rgb8(rs.randint(60, 200, size=(4, 4, 3)).astype(np.uint8))
    s = Scene()
    s.add_light(Rectangle((-1, 1, -1), (-1, 1, -2), (1, 1, -2), Material.light(Spectrum.WHITE())))
    s.add(Rectangle((-1, -1, 0), (1, -1, 0), (1, -1, -2), Material.microfacet(1.0, 1.5, 0.0, False, False, Texture.Image(img), Spectrum.WHITE(), Spectrum.BLACK(), bump_map=bump)))
    prog = s._program(CameraBuilder.new().resolution((16, 16)).build())
    O = oracle_lib.OracleScene(prog)
    ti, tb = _tex_ids(O, P.TEX_IMAGE)[0], _tex_ids(O, P.TEX_BUMP)[0]
    H, W = 3, 5
    lam_u = 0.61
    lam = np.array([oracle_lib.lib().oracle_lambda_sample_one(((lam_u + k / 4.0) % 1.0) if (lam_u + k / 4.0) > 1.0 else lam_u + k / 4.0) for k in range(4)])
    def spec(c, l):
        import ctypes as C
        cf = np.asarray(c, np.float32)
        return oracle_lib.lib().oracle_spectrum_sample(cf.ctypes.data_as(C.POINTER(C.c_float)), C.c_double(l))
    # texel centres return the texel itself (image.rs:99-131: weights 1 / 0)
    for y in range(H):
        for x in range(W):
            uv = np.array([[(x + 0.5) / W, 1.0 - (y + 0.5) / H]])
            got = O.texture_eval(ti, uv, lam_u)[0]
            assert np.allclose(got, [spec(img.data[y, x], l) for l in lam], rtol=1e-12), (x, y)
    # general points: numpy restatement of bilin_interp + value_at
    uv = rs.rand(400, 2)
    got = O.texture_eval(ti, uv, lam_u)
    x = uv[:, 0] * W; y = (1.0 - uv[:, 1]) * H
    xo = np.floor(x - 0.5); yo = np.floor(y - 0.5)
    wx = 1.0 - (x - xo - 0.5); wy = 1.0 - (y - yo - 0.5)
    xi0 = (xo + W).astype(np.int64) % W; yi0 = (yo + H).astype(np.int64) % H; xi1 = (xi0 + 1) % W; yi1 = (yi0 + 1) % H
    S = np.array([[[spec(img.data[yy, xx], l) for l in lam] for xx in range(W)] for yy in range(H)])
    ref = (S[yi0, xi0] * wx[:, None] + S[yi0, xi1] * (1 - wx)[:, None]) * wy[:, None] + (S[yi1, xi0] * wx[:, None] + S[yi1, xi1] * (1 - wx)[:, None]) * (1 - wy)[:, None]
    assert np.allclose(got, ref, rtol=1e-12, atol=1e-15)
    # wrap-around: u = 0 and u = 1 see the same texels
    e = rs.rand(20)
    assert np.allclose(O.texture_eval(ti, np.stack([np.zeros(20), e], -1), lam_u), O.texture_eval(ti, np.stack([np.ones(20), e], -1), lam_u), rtol=1e-12)
    # bump: unit normals; texel centres return the stored normal (image.rs:134-151)
    n = O.bump_eval(tb, uv)
    assert np.allclose(np.linalg.norm(n, axis=1), 1.0, atol=1e-14)
    c = O.bump_eval(tb, np.array([[(1 + 0.5) / 4, 1.0 - (2 + 0.5) / 4]]))[0]
    assert np.allclose(c, bump.data[2, 1], atol=1e-15)
    O.close()


def test_host_blob_carries_textures():
    prog, blob, _ = small_scene("textured")
    B = native.Blob(blob)
    kinds = B.textures["kind"].tolist()
    for k in (P.TEX_SOLID, P.TEX_CHECKER, P.TEX_MARBLE, P.TEX_IMAGE, P.TEX_MANDELBROT, P.TEX_BUMP):
        assert k in kinds
    n_tex = len(kinds)
    for i, t in enumerate(B.textures):
        if t["kind"] == P.TEX_CHECKER: assert t["a"] < i and t["b"] < i
        if t["kind"] == P.TEX_IMAGE: assert int(t["data"]) + int(t["width"]) * int(t["height"]) <= len(B.tex_pixels)
        if t["kind"] == P.TEX_BUMP: assert int(t["data"]) + 3 * int(t["width"]) * int(t["height"]) <= len(B.tex_f64)
        if t["kind"] == P.TEX_MARBLE: assert int(t["data"]) + 1536 <= len(B.tex_f64)
    m = B.materials
    used = [int(v) for col in ("kd_tex", "ks_tex", "tf_tex", "ke_tex", "bump_tex") for v in m[col] if v != 0xFFFFFFFF]
    assert used and max(used) < n_tex
    assert (m["bump_tex"] != 0xFFFFFFFF).sum() >= 3 and (m["ke_tex"] != 0xFFFFFFFF).sum() == 2      # image light + image environment map
    # the Perlin tables in the blob are the oracle's (same Xorshift stream, perlin.rs:31-46)
    import ctypes as C
    tab = np.zeros(1536)
    marble = [t for t in B.textures if t["kind"] == P.TEX_MARBLE]
    oracle_lib.lib().oracle_perlin_tables(C.c_uint64(7), tab.ctypes.data_as(C.POINTER(C.c_double)))
    d = int(marble[0]["data"])
    assert np.array_equal(B.tex_f64[d:d + 1536], tab)
    # a light whose texture has no power (texture.rs:95-101 `unimplemented!()`) is rejected, not rendered wrongly
    s = Scene()
    s.add_light(Rectangle((-1, 1, -1), (-1, 1, -2), (1, 1, -2), Material.light(Texture.Mandelbrot())))
    s.add(Rectangle((-1, -1, 0), (1, -1, 0), (1, -1, -2), Material.diffuse(Spectrum.WHITE())))
    with pytest.raises(RuntimeError, match="power"):
        native.build_blob(s._program(CameraBuilder.new().resolution((8, 8)).build()))


@pytest.mark.gpu
def test_gpu_textured_primary_hits_match_oracle_shading(gpu_ctx):
    """DirectLight at 1 spp with a box filter: every pixel is one camera path whose first-hit albedo comes straight from the
    texture under it; GPU and oracle (same Philox streams) must agree per pixel to rounding except where a path flipped."""
    prog, blob, _ = small_scene("textured", box_filter=True)
    O = oracle_lib.OracleScene(prog)
    G = native.GpuScene(gpu_ctx, blob)
    e = O.render(integrator=1, spp=1, seed=3, rng_mode=1)[0]
    g = G.render(integrator=1, spp=1, seed=3)[0]
    ev, gv = e[..., :3], g[..., :3]
    bad = (np.abs(gv - ev) > 1e-9 * (np.abs(ev) + 1e-6)).any(axis=-1)
    assert bad.mean() <= 0.03, float(bad.mean())
    assert (ev.reshape(-1, 3).std(axis=0) > 0).all()                     # the image is not flat: textures are in play
    G.close(); O.close()
