"""CPU-only: the C-ABI library builds, loads and exports every entry include/paladin_gpu.h declares, and the
product path fails loudly (no CPU fallback) when there is no device."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def pg():
    import paladin_b200 as pg
    if not os.path.exists(pg.lib_path()):
        import __graft_entry__
        __graft_entry__.build()
    return pg


def test_every_declared_symbol_is_exported(pg):
    hdr = open(os.path.join(ROOT, "include", "paladin_gpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(paladin_gpu_\w+)\s*\(", hdr)))
    assert len(names) >= 18
    L = ctypes.CDLL(pg.lib_path())
    for nm in names:
        assert hasattr(L, nm), nm


def test_no_cpu_fallback_without_device(pg):
    if pg.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(pg.PaladinGpuError):
        pg.geev_batch(np.array([2]), np.array([0, 4]), np.array([1, 2, 1, 2]), np.array([1, 1, 2, 2]),
                      np.array([1.0, 2.0, 3.0, 4.0]))
    wr, wi, vr, info = pg.dgeev(np.eye(3))
    assert info < -1000  # the Fortran-signature shim reports the missing device through info


def test_product_never_imports_oracle():
    bad = []
    for d, _, files in os.walk(os.path.join(ROOT, "paladin_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                src = open(os.path.join(d, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle|#include\s+\".*oracle/|scipy", src, re.M):
                    bad.append(f)
    assert not bad, bad
