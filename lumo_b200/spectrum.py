"""Spectrum (sigmoid-polynomial reflectance, Jakob & Hanika 2019), host side.

Mirrors `Spectrum` of the reference (src/tracer/color/spectrum.rs:14-151): `from_rgb`, `from_srgb`, `from_pts`, the
named constants, `sample_one`.  `from_rgb` is the reference's: a trilinear lookup, in f32, in the 64^3 coefficient table
`srgb.coeff` (spectrum/tables.rs:6-84).  That 9.4 MB file is absent from the reference mount, so it is regenerated on
first use by the host library (csrc/host/srgb_table.h: the published optimiser's procedure, ~10 s on 8 cores) and cached
next to this file in the reference's own format ("SPEC", u32 resolution, 64 f32 scale knots, 3 x 64^3 x 3 f32).  Lookups in
the regenerated table reproduce the reference's 33 known-answer triples (spectrum_tests.rs:37-111) within the reference's
own tolerance of 4.6e-4 on all three coefficients (tests/test_spectrum.py).  One node is pinned rather than fitted: full-
brightness white, where the fit is unbounded (any large positive polynomial gives reflectance 1) and the optimiser's
end point is an accident of its rounding — it holds the reference's published answer (spectrum_tests.rs `white_correct`)."""
import ctypes as C
import os
import struct
import numpy as np
from ._tables import TABLES, LAMBDA_MIN, LAMBDA_MAX, Y_INTEGRAL

_N = 95
_RES = 64
_COEFF_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "srgb.coeff")
_WHITE_NODE = (0.001685, -2.276728, 807.041931)      # spectrum_tests.rs: white_correct
_table = None


def _generate_table():
    from . import native
    L = native.host_lib()
    L.lumo_host_srgb_table.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int32, C.c_double, C.c_double]
    L.lumo_host_srgb_table.restype = C.c_int32
    scale = np.zeros(_RES, np.float32); data = np.zeros(3 * _RES ** 3 * 3, np.float32)
    rc = L.lumo_host_srgb_table(scale.ctypes.data_as(C.POINTER(C.c_float)), data.ctypes.data_as(C.POINTER(C.c_float)), os.cpu_count() or 1, 10566.864005283874576, 200.0)
    if rc != 0:
        raise RuntimeError("srgb.coeff generation failed")
    white = (((2 * _RES + (_RES - 1)) * _RES + (_RES - 1)) * _RES + (_RES - 1)) * 3        # from_rgb(1, 1, 1): maxc = 2, last knot, x = y = 1
    data[white:white + 3] = _WHITE_NODE
    tmp = _COEFF_PATH + ".tmp%d" % os.getpid()
    with open(tmp, "wb") as f:
        f.write(b"SPEC"); f.write(struct.pack("<I", _RES)); f.write(scale.tobytes()); f.write(data.tobytes())
    os.replace(tmp, _COEFF_PATH)


def srgb_coeff_table():
    """(scale knots f32[64], coefficients f32[3 * 64^3 * 3]) of srgb.coeff, generated on first use."""
    global _table
    if _table is None:
        if not os.path.exists(_COEFF_PATH) or os.path.getsize(_COEFF_PATH) != 9437448:
            _generate_table()
        raw = np.fromfile(_COEFF_PATH, dtype=np.uint8)
        assert raw[:4].tobytes() == b"SPEC" and int(raw[4:8].view("<u4")[0]) == _RES
        _table = (raw[8:8 + 4 * _RES].view("<f4").copy(), raw[8 + 4 * _RES:].view("<f4").copy())
    return _table


def _table_eval(maxc, xn, yn, zn):
    """tables::srgb::eval (spectrum/tables.rs:31-83), f32 like the reference's TexFloat."""
    f = np.float32
    scale, data = srgb_coeff_table()
    xn, yn, zn = f(xn), f(yn), f(zn)
    x = xn * f(_RES - 1); y = yn * f(_RES - 1)
    xi = min(int(x), _RES - 2); yi = min(int(y), _RES - 2)
    left, right = 0, _RES - 1
    while left < right:
        mid = (left + right) // 2
        if scale[mid] <= zn: left = mid + 1
        else: right = mid
    zi = (left + right) // 2 - 1
    x1 = x - f(xi); x0 = f(1.0) - x1
    y1 = y - f(yi); y0 = f(1.0) - y1
    z1 = (zn - scale[zi]) / (scale[zi + 1] - scale[zi]); z0 = f(1.0) - z1
    dx = 3; dy = _RES * dx; dz = _RES * dy
    offset = (((maxc * _RES + zi) * _RES + yi) * _RES + xi) * 3
    cs = []
    for i in range(3):
        o = offset + i
        x00 = data[o] * x0 + data[o + dx] * x1
        x10 = data[o + dy] * x0 + data[o + dy + dx] * x1
        x01 = data[o + dz] * x0 + data[o + dz + dx] * x1
        x11 = data[o + dz + dy] * x0 + data[o + dz + dy + dx] * x1
        y00 = x00 * y0 + x10 * y1
        y01 = x01 * y0 + x11 * y1
        cs.append(float(y00 * z0 + y01 * z1))
    return cs


class Spectrum:
    """c0, c1, c2, scale — stored as f32 like the reference's `TexFloat` (spectrum.rs:9-19)."""
    __slots__ = ("c0", "c1", "c2", "scale")

    def __init__(self, c0=0.0, c1=0.0, c2=0.0, scale=0.0):
        self.c0, self.c1, self.c2, self.scale = (float(np.float32(v)) for v in (c0, c1, c2, scale))

    @staticmethod
    def from_rgb(r, g, b):                       # spectrum.rs:52-73
        r, g, b = float(r), float(g), float(b)
        c = (r, g, b)
        maxc = 0 if r > g else 1
        maxc = maxc if c[maxc] > b else 2
        if c[maxc] == 0.0 or (r == 0.0 and g == 0.0 and b == 0.0):
            return Spectrum()
        f = np.float32
        mx = f(c[maxc])
        scale = f(2.0) * mx if c[maxc] > 1.0 else f(1.0)
        c0, c1, c2 = _table_eval(maxc, f(c[(maxc + 1) % 3]) / mx, f(c[(maxc + 2) % 3]) / mx, mx / scale)
        return Spectrum(c0, c1, c2, float(scale))

    @staticmethod
    def srgb_decode(v):                          # color/rgb.rs:49-56
        u = v / 255.0
        return u / 12.92 if u <= 0.04045 else ((u + 0.055) / 1.055) ** 2.4

    @staticmethod
    def from_srgb(r, g, b):                      # spectrum.rs:39-42
        return Spectrum.from_rgb(Spectrum.srgb_decode(r), Spectrum.srgb_decode(g), Spectrum.srgb_decode(b))

    @staticmethod
    def from_pts(pts):                           # spectrum.rs:79-95 + dense_spectrum.rs:31-69,99-105
        pairs = sorted(((float(a), float(b)) for a, b in (p.split(":") for p in pts.split())), key=lambda p: p[0])
        lams = np.array([p[0] for p in pairs]); vals = np.array([p[1] for p in pairs])
        dense = np.zeros(_N)
        step = (LAMBDA_MAX - LAMBDA_MIN) / (_N - 1)
        for i in range(_N):
            lam = LAMBDA_MIN + i * step
            b1 = int(np.searchsorted(lams, lam, side="left"))
            if b1 < len(lams) and lams[b1] == lam:
                dense[i] = vals[b1]; continue
            l1, i1 = (lam, 0.0) if b1 == len(lams) else (lams[b1], vals[b1])
            l0, i0 = (lam, 0.0) if b1 == 0 else (lams[b1 - 1], vals[b1 - 1])
            with np.errstate(all="ignore"):
                x1 = (lam - l0) / (l1 - l0)
            dense[i] = (1.0 - x1) * i0 + x1 * i1
        dense = np.nan_to_num(dense, nan=0.0)
        xyz = np.array([np.sum(dense * TABLES[k]) for k in ("X", "Y", "Z")]) / Y_INTEGRAL
        from .color import srgb_from_xyz
        rgb = srgb_from_xyz(xyz)
        return Spectrum.from_rgb(*rgb)

    def sample_one(self, lam):                   # spectrum.rs:108-118 (f32 arithmetic)
        f = np.float32
        l = f(lam)
        x = f(f(f(self.c0) * l) * l) + f(f(self.c1) * l) + f(self.c2)
        return float(f(self.scale) * (f(0.5) + x / (f(2.0) * np.sqrt(f(1.0) + x * x))))

    def is_black(self):
        return self.scale == 0.0

    def __mul__(self, k):                        # spectrum.rs:126-150
        return Spectrum(self.c0, self.c1, self.c2, float(np.float32(self.scale) * np.float32(k)))
    __rmul__ = __mul__

    def as_tuple(self):
        return (self.c0, self.c1, self.c2, self.scale)


def _named(r, g, b):
    return lambda: Spectrum.from_rgb(r, g, b)


# spectrum.rs:23-37 — evaluated lazily (each costs a few ms of Gauss-Newton)
Spectrum.WHITE = _named(1.0, 1.0, 1.0)
Spectrum.BLACK = lambda: Spectrum()
Spectrum.RED = _named(1.0, 0.0, 0.0)
Spectrum.GREEN = _named(0.0, 1.0, 0.0)
Spectrum.BLUE = _named(0.0, 0.0, 1.0)
Spectrum.YELLOW = _named(1.0, 1.0, 0.0)
Spectrum.MAGENTA = _named(1.0, 0.0, 1.0)
Spectrum.CYAN = _named(0.0, 1.0, 1.0)
