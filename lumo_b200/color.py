"""Colour-space constants of the reference evaluated on the host (src/tracer/color/space.rs:50-178,
color/xyz.rs:5-26): XYZ->RGB matrices built from primaries + the D65 white of the 95-sample table,
von Kries white balance, transfer functions.  numpy float64; used for film finalisation and for
filling the camera/film block of the scene blob."""
import numpy as np
from ._tables import TABLES, Y_INTEGRAL


def _to_xyz(table):
    return np.array([np.sum(table * TABLES[k]) / Y_INTEGRAL for k in ("X", "Y", "Z")])


def _to_xyY(xyz):
    s = xyz[0] + xyz[1] + xyz[2]
    return np.array([xyz[0] / s, xyz[1] / s])


def _from_xyY(xy, Y):
    if xy[1] == 0.0:
        return np.zeros(3)
    return np.array([xy[0] * Y / xy[1], Y, (1.0 - xy[0] - xy[1]) * Y / xy[1]])


W_D65 = _from_xyY(_to_xyY(_to_xyz(TABLES["D65"])), 1.0)


def _xyz_to_rgb(r, g, b, W):
    M = np.stack([_from_xyY(np.array(p), 1.0) for p in (r, g, b)]).T
    C = np.linalg.inv(M) @ W
    return np.linalg.inv(M @ np.diag(C))


XYZ_TO_RGB = {
    0: _xyz_to_rgb((0.64, 0.33), (0.3, 0.6), (0.15, 0.06), W_D65),        # sRGB
    1: _xyz_to_rgb((0.68, 0.32), (0.265, 0.69), (0.15, 0.06), W_D65),     # DCI-P3 (default)
    2: _xyz_to_rgb((0.708, 0.292), (0.170, 0.797), (0.131, 0.046), W_D65),  # Rec. 2020
}
XYZ_TO_LMS = np.array([[0.210576, 0.855098, -0.0396983], [-0.417076, 1.177260, 0.0786283], [0.0, 0.0, 0.5168350]])


def srgb_from_xyz(xyz):
    return XYZ_TO_RGB[0] @ np.asarray(xyz)


def wb_matrix(illuminant_name):
    xy = _to_xyY(_to_xyz(TABLES[illuminant_name]))
    diag = (XYZ_TO_LMS @ W_D65) / (XYZ_TO_LMS @ _from_xyY(xy, 1.0))
    return np.linalg.inv(XYZ_TO_LMS) @ np.diag(diag) @ XYZ_TO_LMS


def encode(rgb, color_space):
    """TransferFunction::apply (space.rs:8-36): returns uint8, truncating and saturating."""
    from .native import host_math      # the transfer curve's pow is csrc/common/lumo_math.h: the device kernel's, bit for bit
    c = np.asarray(rgb, dtype=np.float64)
    with np.errstate(invalid="ignore"):
        if color_space == 2:
            beta = 0.018053968510807; alpha = 1.0 + 5.5 * beta
            ec = np.where(c <= beta, 4.5 * c, alpha * host_math("pow", np.maximum(c, 0), 0.45) - (alpha - 1.0))
        else:
            ec = np.where(c <= 0.0031308, 12.92 * c, 1.055 * host_math("pow", np.maximum(c, 0), 1.0 / 2.4) - 0.055)
    v = np.nan_to_num(ec * 255.0, nan=0.0, posinf=255.0, neginf=0.0)
    return np.clip(np.trunc(v), 0, 255).astype(np.uint8)
