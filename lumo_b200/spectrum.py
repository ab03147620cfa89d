"""Spectrum (sigmoid-polynomial reflectance, Jakob & Hanika 2019), host side.

Mirrors `Spectrum` of the reference (src/tracer/color/spectrum.rs:14-151): `from_rgb`, `from_srgb`,
`from_pts`, the named constants, `sample_one`.  The reference evaluates `from_rgb` by trilinear
lookup in a precomputed 64^3 coefficient table (spectrum/tables.rs:6-84, `srgb.coeff`); that 9.4 MB
blob is absent from the reference mount, so the coefficients are found here by running the
Jakob-Hanika Gauss-Newton fit directly for the requested colour (the procedure that generated the
table).  Against the reference's 33 known-answer triples (spectrum_tests.rs:37-111) the fit
agrees to table-interpolation accuracy (tests/test_spectrum.py states the tolerance)."""
import numpy as np
from ._tables import TABLES, LAMBDA_MIN, LAMBDA_MAX, Y_INTEGRAL

_N = 95
_FINE = (_N - 1) * 3 + 1
_XYZ_TO_SRGB = np.array([[3.240479, -1.537150, -0.498535], [-0.969256, 1.875991, 0.041556], [0.055648, -0.204043, 1.057311]])
_SRGB_TO_XYZ = np.array([[0.412453, 0.357580, 0.180423], [0.212671, 0.715160, 0.072169], [0.019334, 0.119193, 0.950227]])


def _interp(data, lam):
    x = (lam - LAMBDA_MIN) * (_N - 1) / (LAMBDA_MAX - LAMBDA_MIN)
    off = np.clip(x.astype(np.int64), 0, _N - 2)
    w = x - off
    return (1.0 - w) * data[off] + w * data[off + 1]


def _init_tables():
    h = (LAMBDA_MAX - LAMBDA_MIN) / (_FINE - 1)
    lam = LAMBDA_MIN + np.arange(_FINE) * h
    xyz = np.stack([_interp(TABLES[k], lam) for k in ("X", "Y", "Z")])
    illum = _interp(TABLES["D65"], lam)
    w = np.full(_FINE, 3.0 / 8.0 * h)
    idx = np.arange(_FINE)
    inner = (idx > 0) & (idx < _FINE - 1)
    w[inner & ((idx - 1) % 3 == 2)] *= 2.0
    w[inner & ((idx - 1) % 3 != 2)] *= 3.0
    illum = illum / np.sum(xyz[1] * illum * w)      # normalise so that white has Y = 1
    rgb_tbl = _XYZ_TO_SRGB @ (xyz * illum * w)
    white = np.sum(xyz * illum * w, axis=1)
    return (lam - LAMBDA_MIN) / (LAMBDA_MAX - LAMBDA_MIN), rgb_tbl, white


_LAM01, _RGB_TBL, _WHITE = _init_tables()


def _lab(rgb):
    xyz = _SRGB_TO_XYZ @ rgb / _WHITE
    d = 6.0 / 29.0
    f = np.where(xyz > d ** 3, np.cbrt(np.maximum(xyz, 0)), xyz / (3 * d * d) + 4.0 / 29.0)
    return np.array([116.0 * f[1] - 16.0, 500.0 * (f[0] - f[1]), 200.0 * (f[1] - f[2])])


def _residual(c, rgb):
    x = (c[0] * _LAM01 + c[1]) * _LAM01 + c[2]
    s = 0.5 * x / np.sqrt(x * x + 1.0) + 0.5
    return _lab(rgb) - _lab(_RGB_TBL @ s)


def _gauss_newton(rgb, c, iters=15):
    eps = 1e-4
    for _ in range(iters):
        r = _residual(c, rgb)
        J = np.zeros((3, 3))
        for i in range(3):
            cp = c.copy(); cp[i] += eps
            cm = c.copy(); cm[i] -= eps
            J[:, i] = (_residual(cp, rgb) - _residual(cm, rgb)) / (2 * eps)
        try:
            c = c - np.linalg.solve(J, r)
        except np.linalg.LinAlgError:
            break
        mx = np.max(np.abs(c))
        if mx > 200.0:
            c = c * (200.0 / mx)
        if np.sum(r * r) < 1e-6:
            break
    return c


def _fit(rgb):
    """coefficients (c0, c1, c2) in nm units for an rgb whose max component is <= 1.
    Follows the table generator's continuation: start at the mid brightness of the same
    chromaticity and walk the brightness towards the target, warm-starting each solve."""
    rgb = np.asarray(rgb, dtype=np.float64)
    mx = rgb.max()
    chroma = rgb / mx
    c = np.zeros(3)
    steps = 24
    start = 0.5
    for k in range(steps + 1):
        z = start + (mx - start) * k / steps
        c = _gauss_newton(chroma * z, c)
    c0n, c1n = LAMBDA_MIN, 1.0 / (LAMBDA_MAX - LAMBDA_MIN)
    A, B, C = c
    return (A * c1n * c1n, B * c1n - 2 * A * c0n * c1n * c1n, C - B * c0n * c1n + A * (c0n * c1n) ** 2)


class Spectrum:
    """c0, c1, c2, scale — stored as f32 like the reference's `TexFloat` (spectrum.rs:9-19)."""
    __slots__ = ("c0", "c1", "c2", "scale")
    _cache = {}

    def __init__(self, c0=0.0, c1=0.0, c2=0.0, scale=0.0):
        self.c0, self.c1, self.c2, self.scale = (float(np.float32(v)) for v in (c0, c1, c2, scale))

    @staticmethod
    def from_rgb(r, g, b):                       # spectrum.rs:52-73
        r, g, b = float(r), float(g), float(b)
        mx = max(r, g, b)
        if mx == 0.0:
            return Spectrum()
        key = (r, g, b)
        if key not in Spectrum._cache:
            scale = 2.0 * mx if mx > 1.0 else 1.0
            c = _fit(np.array([r, g, b]) / scale)
            Spectrum._cache[key] = (c[0], c[1], c[2], scale)
        return Spectrum(*Spectrum._cache[key])

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
