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


def test_every_tuning_option_round_trips_without_a_device(lib):
    """Every field of `Options` (csrc/common.cuh) is reachable by name through set_option / get_option -- the probes
    and the variant tests depend on it -- and the round-2 defaults are the measured ones; unknown names are refused."""
    from lair_b200 import _ffi
    from lair_b200._ffi import LairB200Error
    text = open(os.path.join(ROOT, "lair_b200", "csrc", "common.cuh")).read()
    body = text[text.index("struct Options"):]
    body = body[:body.index("};")]
    names = re.findall(r"int64_t\s+(\w+)\s*=", body)
    assert len(names) >= 30
    for name in names:
        v = _ffi.get_option(name)
        _ffi.set_option(name, v)
        assert _ffi.get_option(name) == v, name
    assert _ffi.get_option("panel_cluster") == 3 and _ffi.get_option("trsm_strip") == 2
    assert _ffi.get_option("pair_k512") == 16384 and _ffi.get_option("drain_rows") == 1 and _ffi.get_option("pair_small") == 0
    with pytest.raises(LairB200Error):
        _ffi.set_option("no_such_option", 1)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under lair_b200/ may reference it."""
    pkg = os.path.join(ROOT, "lair_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", text, re.M), os.path.join(dirpath, f)
                assert "lair_oracle" not in text or f.endswith((".cu", ".cuh")) and "oracle/lair_oracle.hpp" in text, os.path.join(dirpath, f)


def _split_params(params: str):
    params = params.strip()
    return [] if params in ("", "void") else [p.strip() for p in params.split(",")]


def test_rust_shim_declarations_match_header():
    """rust/src/ffi.rs (source-only: no Rust toolchain in the image) binds only symbols the header
    declares, with the same number of parameters and the same pointer-ness per parameter."""
    header = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "lair_b200.h")).read(), flags=re.S)
    c_decl = {m.group(1): _split_params(m.group(2))
              for m in re.finditer(r"LAIR_B200_API\s+[\w\s\*]+?\b(lair_b200_\w+)\s*\(([^)]*)\)", header)}
    rust = re.sub(r"//.*", "", open(os.path.join(ROOT, "rust", "src", "ffi.rs")).read())
    block = rust[rust.index('extern "C"'):]
    r_decl = {m.group(1): _split_params(m.group(2))
              for m in re.finditer(r"pub fn (lair_b200_\w+)\s*\(([^)]*)\)", block, flags=re.S)}
    assert len(r_decl) >= 40
    for name, r_params in r_decl.items():
        assert name in c_decl, f"{name} is not declared in include/lair_b200.h"
        c_params = c_decl[name]
        assert len(r_params) == len(c_params), (name, r_params, c_params)
        for rp, cp in zip(r_params, c_params):
            cp = cp.replace("lair_b200_lu_t*", "void**").replace("lair_b200_lu_t", "void*")  # opaque handle typedef
            assert ("*" in rp) == ("*" in cp), (name, rp, cp)
            if "*" in cp:
                assert ("*const" in rp) == cp.startswith("const "), (name, rp, cp)
    # the drop-in rows of SURVEY 8b are all bound
    for p in "sdcz":
        assert f"lair_b200_{p}getrf" in r_decl and f"lair_b200_{p}getrs" in r_decl
    for p in "sd":
        assert f"lair_b200_{p}gesv" in r_decl and f"lair_b200_{p}getrf_batched" in r_decl
