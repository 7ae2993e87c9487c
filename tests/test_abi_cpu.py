"""CPU-side checks of the boundary: the library builds/loads, exports exactly what
include/astroburst_cuda.h declares, and the host wrappers validate arguments (no compute)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "astroburst_cuda.h")


@pytest.fixture(scope="module")
def so_path():
    from astroburst_b200 import build

    return build.build()


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ab_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(so_path):
    out = subprocess.check_output(["nm", "-D", "--defined-only", so_path], text=True)
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    declared = _declared_symbols()
    assert declared, "no declarations parsed"
    missing = [s for s in declared if s not in exported]
    assert not missing, f"declared in the header but not exported: {missing}"
    extra = sorted(s for s in exported if s.startswith("ab_") and s not in declared)
    assert not extra, f"exported but undeclared: {extra}"


def test_ctypes_table_matches_header(so_path):
    from astroburst_b200 import _lib

    assert sorted(_lib.SIGNATURES) == _declared_symbols()
    lib = _lib.lib()
    assert lib.ab_abi_version() == 1


def test_no_cpu_fallback_without_device(so_path):
    import torch
    import astroburst_b200 as ab

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(ab.AstroBurstError):
        ab.Context()
    with pytest.raises(ab.AstroBurstError):
        import numpy as np
        ab.stack_images([np.zeros((4, 4), np.float32)], ab.StackConfig(align=False))


def test_product_never_imports_oracle():
    """Only tests/, bench.py (cpu_baseline / --impl reference) and smoke() may touch oracle/."""
    pkg = os.path.join(ROOT, "astroburst_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh", ".cpp", ".hpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower() or f == "__init__.py" and "oracle" not in text, (dirpath, f)
    for f in os.listdir(os.path.join(ROOT, "include")):
        p = os.path.join(ROOT, "include", f)
        if os.path.isfile(p):
            assert "liboracle" not in open(p).read()
