"""CPU: the C-ABI library builds, loads, exports every symbol include/lair_b200.h declares, and
fails loudly without a device (no compute calls here)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "lair_b200.h")).read()
    return sorted(set(re.findall(r"LAIR_B200_API\s+[\w\s\*]+?\b(lair_b200_\w+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from lair_b200 import build, _ffi
    build.build()
    return _ffi.lib()


def test_header_declares_expected_surface():
    syms = _header_symbols()
    for p in "sdcz":
        assert f"lair_b200_{p}getrf" in syms and f"lair_b200_{p}getrs" in syms and f"lair_b200_{p}getrf_dev" in syms
    for p in "sd":
        for name in ("gesv", "getrf_batched", "getrf_dev", "getrs_dev", "getrf_batched_dev", "laswp_dev", "trsm_dev",
                     "gemm_minus_dev"):
            assert f"lair_b200_{p}{name}" in syms
    assert len(syms) >= 35


def test_library_exports_every_declared_symbol(lib):
    missing = [s for s in _header_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_ctypes_binding_matches_header(lib):
    from lair_b200 import _ffi
    assert sorted(_ffi.EXPORTED_SYMBOLS) == _header_symbols()


def test_version_and_error_string(lib):
    assert lib.lair_b200_version() >= 100
    assert isinstance(lib.lair_b200_last_error(), bytes)


def test_fails_loudly_without_device(lib):
    """No CPU fallback: on a box without a GPU every compute entry point returns NO_DEVICE."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import lair_b200
    from lair_b200._ffi import LairB200Error, ERR_NO_DEVICE
    with pytest.raises(LairB200Error) as e:
        lair_b200.lapack.getrf(np.eye(3))
    assert e.value.status == ERR_NO_DEVICE
    with pytest.raises(LairB200Error):
        lair_b200.lapack.getrf_batched(np.zeros((2, 32, 32)))
    with pytest.raises(LairB200Error):
        lair_b200.equation.solve(np.eye(2), np.ones(2))


def test_host_layer_shape_errors_need_no_device():
    """Shape validation mirrors the reference and happens before any device work."""
    import lair_b200
    from lair_b200 import InvalidInput
    with pytest.raises(InvalidInput.Shape, match="input matrix is not square"):
        lair_b200.equation.solve(np.zeros((2, 3)), np.zeros(2))
    with pytest.raises(InvalidInput.Shape, match="must be the same as the number of rows"):
        lair_b200.equation.solve(np.eye(3), np.zeros(2))
    with pytest.raises(AssertionError):
        lair_b200.lapack.getrs(np.eye(3), [0, 1], np.zeros(3))
    assert str(InvalidInput.Shape("x")) == "shape error: x" and str(InvalidInput.Value("y")) == "value error: y"


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under lair_b200/ may reference it."""
    pkg = os.path.join(ROOT, "lair_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", text, re.M), os.path.join(dirpath, f)
                assert "lair_oracle" not in text or f.endswith((".cu", ".cuh")) and "oracle/lair_oracle.hpp" in text, os.path.join(dirpath, f)
