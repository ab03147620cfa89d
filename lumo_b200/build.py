"""In-tree native build: liblumo_host.so (C++ scene builder) and liblumo_gpu.so (CUDA kernels +
the C ABI of include/lumo_gpu.h) for sm_100a.  `python -m lumo_b200.build [host|gpu|all]`."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
HOST_SO = os.path.join(HERE, "liblumo_host.so")
GPU_SO = os.path.join(HERE, "liblumo_gpu.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _sources(sub, exts):
    out = []
    for d in (os.path.join(CSRC, sub), os.path.join(CSRC, "common")):
        for f in sorted(os.listdir(d)):
            if f.endswith(exts):
                out.append(os.path.join(d, f))
    return out


def build_host(force=False, verbose=False):
    srcs = _sources("host", (".cpp", ".h"))
    if force or _stale(HOST_SO, srcs):
        cmd = ["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-pthread", "-Wall", "-Wno-unused-function",
               "-o", HOST_SO] + [s for s in srcs if s.endswith(".cpp")]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return HOST_SO


def build_gpu(force=False, verbose=False):
    srcs = _sources("gpu", (".cu", ".cuh", ".h"))
    include = os.path.join(os.path.dirname(HERE), "include")
    srcs_all = srcs + [os.path.join(include, f) for f in os.listdir(include)]
    if force or _stale(GPU_SO, srcs_all):
        # -fmad=false: the reference never contracts a*b+c; parity of t / barycentrics / ids depends on it.
        cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-prec-div=true", "-prec-sqrt=true",
               "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "-I", include, "-o", GPU_SO] + [s for s in srcs if s.endswith(".cu")]
        if os.environ.get("LUMO_PTXAS_V"):
            cmd[1:1] = ["-Xptxas", "-v"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return GPU_SO


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("host", "all"):
        print(build_host(force=True, verbose=True))
    if what in ("gpu", "all"):
        print(build_gpu(force=True, verbose=True))
