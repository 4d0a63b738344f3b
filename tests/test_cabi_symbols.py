"""CPU-side checks of the C-ABI boundary: the library loads, exports every symbol include/gpb200.h declares, and refuses
to compute without a GPU (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

from gaussian_process_optimization_b200 import _lib, native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "gpb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gpb_[a-z_A-Z0-9]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libgpb200.so does not export %s" % n
        assert n in _lib.SIGNATURES, "python binding missing for %s" % n
    assert sorted(_lib.SIGNATURES) == names


def test_version_and_workspace_query():
    lib = _lib.load()
    assert lib.gpb_version() >= 100
    small = lib.gpb_model_workspace_bytes(100, 2, 1, 128)
    big = lib.gpb_model_workspace_bytes(16384, 16, 1, 4096)
    assert 0 < small < big
    assert big > 3 * 16384 * 16384 * 8          # L, L^-1, Ky^-1
    assert big < 12 * 1024 ** 3


@pytest.mark.skipif(_lib.load().gpb_device_count() > 0, reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback():
    with pytest.raises(_lib.GpbError):
        native.kern_K("rbf", np.zeros((4, 2)), None, 1.0, [1.0, 1.0])
    with pytest.raises(_lib.GpbError):
        native.NativeModel("rbf", True, 2)
