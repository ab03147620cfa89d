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


def test_rust_crate_declares_every_export():
    """rust/lumo-gpu-sys/src/lib.rs (the FFI crate a lumo maintainer links; no Rust toolchain exists here, so it is source
    only) declares every function of include/lumo_gpu.h with the same number of arguments, and the two structs field for field."""
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "lumo_gpu.h")).read(), flags=re.S)
    rs = re.sub(r"//.*", "", open(os.path.join(ROOT, "rust", "lumo-gpu-sys", "src", "lib.rs")).read())
    for name in _declared():
        m = re.search(r"\b%s\s*\(([^;]*?)\)\s*;" % name, hdr, flags=re.S)
        r = re.search(r"pub fn %s\s*\(([^;]*?)\)\s*(?:->[^;]*)?;" % name, rs, flags=re.S)
        assert r, "lib.rs lacks " + name
        c_args = [a for a in m.group(1).split(",") if a.strip() and a.strip() != "void"]
        r_args = [a for a in r.group(1).split(",") if a.strip()]
        assert len(c_args) == len(r_args), (name, c_args, r_args)
    # struct fields in order
    def c_fields(struct):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (struct, struct), hdr, flags=re.S).group(1)
        out = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl: continue
            names = decl.split(None, 1)[1] if " " in decl else decl
            for n in names.split(","):
                out.append(re.sub(r"[\*\s]|\[.*\]", "", n))
        return out
    def rs_fields(struct):
        body = re.search(r"pub struct %s \{(.*?)\}" % struct, rs, flags=re.S).group(1)
        return re.findall(r"pub (\w+)\s*:", body)
    for st in ("lumo_render_params", "lumo_film_accum"):
        assert c_fields(st) == rs_fields(st), (st, c_fields(st), rs_fields(st))
    assert "counters: [u64; 10]" in rs and "uint64_t counters[10]" in hdr
