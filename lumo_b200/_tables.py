"""Spectral data tables (CIE 1931 observer, illuminants, refractive indices; 95 samples, 5 nm,
360..830 nm) parsed from csrc/host/spectra_data.h, which tools/gen_spectra_data.py extracted as
data from the reference's src/tracer/color/samples.rs:15-276."""
import os
import re
import numpy as np

_HDR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "host", "spectra_data.h")


def _load():
    src = open(_HDR).read()
    out = {}
    for m in re.finditer(r"static const double (\w+)\[95\] = \{([^}]*)\}", src):
        vals = [float(v) for v in m.group(2).replace("\n", " ").split(",") if v.strip()]
        assert len(vals) == 95
        out[m.group(1)] = np.array(vals, dtype=np.float64)
    return out


TABLES = _load()
ILLUMINANTS = ["A", "D50", "D65", "F2", "F7", "CORNELL"]   # ids used by the scene program
LAMBDA_MIN, LAMBDA_MAX = 360.0, 830.0
Y_INTEGRAL = 106.856895   # src/tracer/color/xyz.rs:32
