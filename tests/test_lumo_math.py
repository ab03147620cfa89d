"""csrc/common/lumo_math.h — the FMA-free sin / cos / atan2 / acos / atanh / cosh / exp / log / pow that the device kernels,
the host builder and the oracle all compile from one source (so that a sampled direction has the same bits on every side).
CPU: accuracy against the host libm and g++-vs-g++ agreement of the two host builds.  GPU: the nvcc build returns the
same bits as the g++ builds on every input."""
import ctypes as C
import numpy as np
import pytest
import oracle_lib
from lumo_b200 import native

CASES = {   # name -> (numpy reference, sampler of the argument range the renderer uses, second argument or None, ulp bound vs libm)
    "sin": (np.sin, lambda rs, n: np.concatenate([rs.uniform(-7, 7, n), rs.uniform(-400, 400, n), rs.uniform(-1e-6, 1e-6, n)]), None, 1),
    "cos": (np.cos, lambda rs, n: np.concatenate([rs.uniform(-7, 7, n), rs.uniform(-400, 400, n), rs.uniform(-1e-6, 1e-6, n)]), None, 1),
    "atan2": (np.arctan2, lambda rs, n: rs.uniform(-3, 3, 3 * n), "uniform", 2),
    "acos": (np.arccos, lambda rs, n: np.concatenate([rs.uniform(-1, 1, 2 * n), 1 - rs.uniform(0, 1e-6, n)]), None, 1),
    "atanh": (np.arctanh, lambda rs, n: np.concatenate([rs.uniform(-0.98, 0.86, 2 * n), rs.uniform(-1e-4, 1e-4, n)]), None, 4),
    "cosh": (np.cosh, lambda rs, n: rs.uniform(-3, 3, 3 * n), None, 2),
    "exp": (np.exp, lambda rs, n: np.concatenate([rs.uniform(-30, 30, 2 * n), rs.uniform(-700, 700, n)]), None, 1),
    "log": (np.log, lambda rs, n: np.concatenate([rs.uniform(1e-9, 10, 2 * n), np.exp(rs.uniform(-300, 300, n))]), None, 1),
    "pow": (np.power, lambda rs, n: rs.uniform(1e-4, 2.0, 3 * n), 1.0 / 2.4, 6),
}
SPECIAL = np.array([0.0, -0.0, 1.0, -1.0, 0.5, np.pi, -np.pi, np.pi / 2, np.pi / 4, 1e-300, 1e300, np.inf, -np.inf, np.nan, 0.78539816339744839, 2.0 ** -30])


def _inputs(name, n=200000):
    ref, gen, second, bound = CASES[name]
    rs = np.random.RandomState(abs(hash(name)) % (2 ** 31))
    x = np.concatenate([gen(rs, n), SPECIAL])
    y = None
    if second == "uniform": y = np.concatenate([rs.uniform(-3, 3, len(x) - len(SPECIAL)), SPECIAL[::-1]])
    elif second is not None: y = np.full(len(x), second)
    return x, y


def _oracle_eval(name, x, y):
    L = oracle_lib.lib()
    out = np.empty_like(x)
    L.oracle_math_eval(C.c_int(native.MATH_FN[name]), x.ctypes.data_as(C.POINTER(C.c_double)), None if y is None else y.ctypes.data_as(C.POINTER(C.c_double)),
                       C.c_uint64(len(x)), out.ctypes.data_as(C.POINTER(C.c_double)))
    return out


def _ulps(a, b):
    with np.errstate(invalid="ignore", over="ignore"):
        u = np.abs(np.nextafter(b, np.inf) - b)
        e = np.abs(a - b) / u
    e[(a == b) | (np.isnan(a) & np.isnan(b))] = 0.0
    return e


@pytest.mark.parametrize("name", list(CASES))
def test_accuracy_against_libm_and_host_builds_agree(name):
    ref, _, _, bound = CASES[name]
    x, y = _inputs(name)
    got = _oracle_eval(name, x, y)
    host = native.host_math(name, x, y)
    assert np.array_equal(got.view(np.uint64), host.view(np.uint64)), "oracle and host builds of lumo_math.h differ"
    with np.errstate(all="ignore"):
        want = ref(x) if y is None else ref(x, y)
    ok = np.isfinite(want) & np.isfinite(x) & (np.abs(want) > 1e-290)
    if name in ("sin", "cos"): ok &= np.abs(x) < 1e5
    if name == "pow": ok &= (x >= 1e-6) & (x <= 100.0)      # the film's range; exp(y log x) loses bits for huge |y log x|
    e = _ulps(got[ok], want[ok])
    assert e.max() <= bound, (name, float(e.max()), float(x[ok][e.argmax()]))


def test_sincos_pair_and_film_pow_are_the_same_bits():
    """lm_sincos (used where the kernels need both) returns lm_sin / lm_cos; color.encode takes its pow from the same source."""
    from lumo_b200 import color
    x = np.linspace(0.0, 1.2, 4001)
    assert np.array_equal(color.encode(x, 1), np.clip(np.trunc(np.where(x <= 0.0031308, 12.92 * x, 1.055 * native.host_math("pow", x, 1 / 2.4) - 0.055) * 255.0), 0, 255).astype(np.uint8))


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_device_bits_equal_host_bits(gpu_ctx, name):
    """nvcc -fmad=false and g++ -ffp-contract=off compile lumo_math.h to the same function: every input, every bit."""
    x, y = _inputs(name, n=400000)
    dev = gpu_ctx.math_eval(name, x, y)
    host = _oracle_eval(name, x, y)
    both_nan = np.isnan(dev) & np.isnan(host)
    same = (dev.view(np.uint64) == host.view(np.uint64)) | both_nan
    assert same.all(), (name, int((~same).sum()), x[~same][:5], dev[~same][:5], host[~same][:5])

