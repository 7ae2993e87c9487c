"""Build libastroburst_cuda.so in-tree with nvcc for sm_100a.

    python -m astroburst_b200.build [--force]

The library is written to astroburst_b200/lib/ (git-ignored, travels to the GPU box with the
snapshot).  nvcc cross-compiles without a GPU.  Sources are compiled in parallel, one object
per .cu file, then linked with -shared.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(LIBDIR, "obj")
SO = os.path.join(LIBDIR, "libastroburst_cuda.so")

NVCC = os.environ.get("AB_NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",               # Rust never contracts a*b+c; parity needs separate roundings
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC,-O3",
    "-ccbin", "/usr/bin/g++",
    "-I", os.path.join(ROOT, "include"),
    "-I", CSRC,
]


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    hdrs.append(os.path.join(ROOT, "include", "astroburst_cuda.h"))
    return hdrs


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src: str, force: bool) -> str:
    obj = os.path.join(OBJDIR, os.path.basename(src)[:-3] + ".o")
    if force or _stale(obj, [src] + _deps()):
        cmd = [NVCC] + NVCC_FLAGS + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj


def build_variant(name: str, defines) -> str:
    """Developer A/B builds: the same sources with extra -D flags -> lib/variants/<name>.so
    (select it with AB_LIBRARY=<path>)."""
    vdir = os.path.join(LIBDIR, "variants")
    odir = os.path.join(vdir, name + "_obj")
    os.makedirs(odir, exist_ok=True)
    flags = NVCC_FLAGS + [f"-D{d}" for d in defines]

    def one(src):
        obj = os.path.join(odir, os.path.basename(src)[:-3] + ".o")
        r = subprocess.run([NVCC] + flags + ["-c", src, "-o", obj], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return obj

    srcs = _sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(one, srcs))
    so = os.path.join(vdir, name + ".so")
    r = subprocess.run([NVCC, "-shared", "-o", so] + objs +
                       ["-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", "/usr/bin/g++",
                        "-lcudart_static", "-lcufft"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    for o in objs:
        os.remove(o)
    return so


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJDIR, exist_ok=True)
    srcs = _sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, force), srcs))
    if force or _stale(SO, objs):
        cmd = [NVCC, "-shared", "-o", SO] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                     "-ccbin", "/usr/bin/g++", "-lcudart_static", "-lcufft"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if verbose:
        print(SO)
    return SO


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--variant":  # --variant NAME DEF[=VAL] ...
        print(build_variant(sys.argv[2], sys.argv[3:]))
    else:
        build(force="--force" in sys.argv, verbose=True)
