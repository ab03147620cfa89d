"""The C-ABI library loads without a GPU and exports every symbol include/lumo_gpu.h declares."""
import ctypes
import os
import re
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "lumo_gpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lumo_gpu_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from lumo_b200 import build
    so = build.build_gpu()
    L = ctypes.CDLL(so)
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), "missing export: " + n


def test_no_gpu_is_an_error_code_not_a_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from lumo_b200 import native
    with pytest.raises(RuntimeError):
        native.GpuContext(0)                          # status code + message; nothing renders on the CPU
    L = native.gpu_lib()
    assert L.lumo_gpu_last_error().decode() != ""


def test_product_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/."""
    pkg = os.path.join(ROOT, "lumo_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle_lib" not in txt and "liblumo_oracle" not in txt and "oracle/" not in txt, os.path.join(dp, f)
