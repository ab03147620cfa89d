"""Host mirror of the reference's image ingest (src/image.rs): PNG decode to 8-bit RGB triples, `Image<Spectrum>` for
albedo / emission maps, `Image<Normal>` for bump maps, Radiance RGBE for environment maps.  The arrays built here are
shipped inside the scene program (program.TAG_TEXTURE) and end up in the device blob (LSEC_TEX_PIXELS / LSEC_TEX_F64);
the bilinear fetch itself runs on the device (csrc/gpu/shade.cuh, image.rs:99-185).

No `png` crate here: the decoder below handles what image.rs:19-79 handles — 8-bit grey, grey+alpha, RGB, RGBA and
1/2/4/8-bit palettes, non-interlaced — with the standard filters, via zlib.
"""
import struct
import zlib
import numpy as np
from .spectrum import Spectrum


def decode_png(data):
    """bytes -> uint8 array [H, W, 3] (image.rs:19-79: alpha dropped, grey replicated, palette expanded)."""
    if data[:8] != b"\x89PNG\r\n\x1a\n":
        raise ValueError("not a PNG file")
    pos, idat, plte, hdr = 8, [], None, None
    while pos < len(data):
        n, tag = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        pos += 12 + n
        if tag == b"IHDR": hdr = struct.unpack(">IIBBBBB", body)
        elif tag == b"PLTE": plte = np.frombuffer(body, np.uint8).reshape(-1, 3)
        elif tag == b"IDAT": idat.append(body)
        elif tag == b"IEND": break
    w, h, depth, ctype, _, _, interlace = hdr
    if interlace: raise ValueError("interlaced PNG not supported")
    chans = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[ctype]
    if depth == 16 or (depth < 8 and ctype != 3 and ctype != 0): raise ValueError("unsupported PNG bit depth")
    bpp = max(1, chans * depth // 8)
    stride = (w * chans * depth + 7) // 8
    raw = zlib.decompress(b"".join(idat))
    out = np.zeros((h, stride), np.uint8)
    prev = np.zeros(stride, np.int32)
    p = 0
    for y in range(h):
        ft = raw[p]; line = np.frombuffer(raw, np.uint8, stride, p + 1).astype(np.int32); p += 1 + stride
        if ft == 0: cur = line
        elif ft == 2: cur = (line + prev) & 255
        else:
            cur = np.zeros(stride, np.int32)
            for i in range(stride):
                a = cur[i - bpp] if i >= bpp else 0
                b = prev[i]
                c = prev[i - bpp] if i >= bpp else 0
                if ft == 1: pred = a
                elif ft == 3: pred = (a + b) >> 1
                else:
                    pa, pb, pc = abs(b - c), abs(a - c), abs(a + b - 2 * c)
                    pred = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
                cur[i] = (line[i] + pred) & 255
        out[y] = cur; prev = cur
    if ctype == 3:                                   # image.rs:27-51: the pixel index runs over the whole byte buffer
        per = 8 // depth                             # (rows are not re-aligned) and sub-byte pixels count from the LOW bits
        gi = np.arange(h * w)
        byte = out.reshape(-1)[gi // per].astype(np.int32)
        pidx = (byte >> (depth * (gi % per))) & ((1 << depth) - 1)
        return plte[pidx].reshape(h, w, 3)
    if ctype in (0, 4):                              # image.rs:52-59
        if depth < 8: raise ValueError("unsupported PNG bit depth")
        g = out.reshape(h, w, chans)[:, :, 0]
        return np.stack([g, g, g], axis=-1)
    return np.ascontiguousarray(out.reshape(h, w, chans)[:, :, :3])   # image.rs:60-72


def encode_png(rgb):
    """uint8 [H, W, 3] -> PNG bytes (filter 0); used by the tests and `Film.save`-style output."""
    rgb = np.ascontiguousarray(rgb, np.uint8)
    h, w, _ = rgb.shape
    raw = b"".join(b"\x00" + rgb[y].tobytes() for y in range(h))
    def chunk(tag, body): return struct.pack(">I", len(body)) + tag + body + struct.pack(">I", zlib.crc32(tag + body) & 0xFFFFFFFF)
    return b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(raw)) + chunk(b"IEND", b"")


def rgbe_to_rgb(r, g, b, e):                         # color/rgb.rs:78-92
    if e == 0: return (0.0, 0.0, 0.0)
    v = 2.0 ** (int(e) - 128) / 256.0
    return (0.5 + v * r, 0.5 + v * g, 0.5 + v * b)


class Image:
    """`Image<Spectrum>` (kind 'spectrum': data [H, W, 4] of c0, c1, c2, scale; mean = Texture::power) or
    `Image<Normal>` (kind 'normal': data [H, W, 3] unit normals)."""
    def __init__(self, kind, data, mean=None):
        self.kind, self.data, self.mean = kind, np.ascontiguousarray(data, np.float64), mean
        self.height, self.width = self.data.shape[:2]

    @staticmethod
    def _spectra(rgb_lin):
        """per-pixel Spectrum::from_rgb, one fit per distinct colour (the reference looks the coefficients up in
        srgb.coeff, which is not in the mount: lumo_b200/spectrum.py fits them)."""
        h, w, _ = rgb_lin.shape
        flat = rgb_lin.reshape(-1, 3)
        uniq, inv = np.unique(flat, axis=0, return_inverse=True)
        coef = np.array([Spectrum.from_rgb(*c).as_tuple() for c in uniq], np.float64).reshape(-1, 4)
        return coef[inv.reshape(-1)].reshape(h, w, 4)

    @staticmethod
    def from_rgb8(rgb):                              # image.rs:257-277
        rgb = np.asarray(rgb, np.uint8)
        lut = np.array([Spectrum.srgb_decode(v) for v in range(256)])
        lin = lut[rgb]
        s = np.zeros(3)
        for px in lin.reshape(-1, 3): s = s + px     # fold in pixel order
        mean = Spectrum.from_rgb(*(s / float(lin.shape[0] * lin.shape[1])))
        return Image("spectrum", Image._spectra(lin), mean)

    @staticmethod
    def from_png(data): return Image.from_rgb8(decode_png(data))

    @staticmethod
    def bump_from_rgb8(rgb):                         # image.rs:154-180
        v = np.asarray(rgb, np.uint8).astype(np.float64) / 128.0 - 1.0
        ln = np.sqrt(np.maximum(v[..., 0] * v[..., 0] + v[..., 1] * v[..., 1] + v[..., 2] * v[..., 2], 0.0))
        return Image("normal", v / ln[..., None])

    @staticmethod
    def bump_from_png(data): return Image.bump_from_rgb8(decode_png(data))

    @staticmethod
    def from_hdri_bytes(data):                       # image.rs:204-247 (flat RGBE only, as the reference reads it)
        lines = data.split(b"\n")
        assert lines[0].strip() == b"#?RADIANCE"
        pos = len(lines[0]) + 1
        for ln in lines[1:]:
            pos += len(ln) + 1
            if ln[:1] in (b"+", b"-"):
                f = ln.split()
                assert f[0][:1] == b"-" and f[2][:1] == b"+"
                h, w = int(f[1]), int(f[3])
                break
        px = np.frombuffer(data, np.uint8, w * h * 4, pos).reshape(h, w, 4)
        lin = np.array([rgbe_to_rgb(*p) for p in px.reshape(-1, 4)]).reshape(h, w, 3)
        s = np.zeros(3)
        for p in lin.reshape(-1, 3): s = s + p
        mean = Spectrum.from_rgb(*(s / float(w * h)))
        return Image("spectrum", Image._spectra(lin), mean)

    @staticmethod
    def mean_vec3_from_rgb8(rgb):                    # image.rs:82-97
        rgb = np.asarray(rgb, np.uint8).reshape(-1, 3)
        scale = 1.0 / float(rgb.shape[0])
        acc = np.zeros(3)
        for p in rgb: acc = acc + scale * p.astype(np.float64) / 256.0
        return acc
