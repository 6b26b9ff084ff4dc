"""The C-ABI library loads and exports every symbol include/hgsfa.h declares; without a GPU every
compute entry point fails loudly (there is no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from pyfaceanalysis_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "hgsfa.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hgsfa_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 18 and "hgsfa_plan_execute" in names and "hgsfa_gauss_regress" in names
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert getattr(raw, n) is not None, n
    bound = {s[0] for s in _lib.SYMBOLS}
    assert set(names) == bound, (set(names) ^ bound)
    assert lib.hgsfa_version() == 1


def test_constants_match_header():
    src = open(os.path.join(ROOT, "include", "hgsfa.h")).read()
    assert "HGSFA_TILE 128" in src and _lib.TILE == 128
    assert "HGSFA_U8 = 0, HGSFA_F32 = 1, HGSFA_F64 = 2" in src and (_lib.U8, _lib.F32, _lib.F64) == (0, 1, 2)
    assert "HGSFA_NEAREST = 0, HGSFA_BILINEAR = 2" in src and (_lib.NEAREST, _lib.BILINEAR) == (0, 2)


def _has_gpu():
    try:
        return _lib.device_count() > 0
    except _lib.HgsfaError:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the behaviour of a box WITHOUT a CUDA device")
def test_no_cpu_fallback(tiny_flow, classifiers):
    from pyfaceanalysis_b200 import GpuFlow, GpuGaussianClassifier, extract_subimages
    with pytest.raises(_lib.HgsfaError):
        GpuFlow(tiny_flow)
    with pytest.raises(_lib.HgsfaError):
        GpuGaussianClassifier(classifiers[0])
    with pytest.raises(_lib.HgsfaError):
        extract_subimages(np.zeros((8, 8), dtype=np.uint8), np.array([[0.0, 0.0, 7.0, 7.0]]))


def test_bad_blob_is_rejected_before_touching_the_device():
    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.hgsfa_plan_create(b"x" * 10, 10, 0, ctypes.byref(h)) != 0
    assert b"too small" in lib.hgsfa_last_error()
    assert lib.hgsfa_plan_create(None, 0, 0, ctypes.byref(h)) != 0
