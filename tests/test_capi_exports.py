"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/*.h declares.
No compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from adsurf_b200 import _capi
    return _capi.lib()


def declared_functions():
    text = open(os.path.join(ROOT, "include", "adsurf_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(adsurf_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    names = declared_functions()
    assert len(names) >= 10
    for name in names:
        assert hasattr(lib, name), name
    from adsurf_b200 import _capi
    assert sorted(_capi.EXPORTS) == names


def test_version_and_error_strings(lib):
    assert b"sm_100a" in lib.adsurf_version()
    assert lib.adsurf_error_string(0) == b"ok"
    for code in range(-6, 0):
        assert lib.adsurf_error_string(code) != b"unknown error"


def test_workspace_sizes_are_host_side(lib):
    from adsurf_b200 import _capi
    n = lib.adsurf_workspace_bytes(_capi.OP_GRID_BWD, 4096, 20, 128, 1024)
    assert 0 < n < (1 << 30)
    assert lib.adsurf_workspace_bytes(_capi.OP_POINTS_FWD, 10, 6, 100, 0) == 0
    assert lib.adsurf_workspace_bytes(_capi.OP_DMF_STEP, 100, 3, 174, 200) > 0
    assert lib.adsurf_workspace_bytes(_capi.OP_GRID_BWD, 0, 20, 128, 1024) == 0


def test_argument_validation_without_gpu(lib):
    """NULL pointers and bad shapes are rejected before any CUDA call"""
    assert lib.adsurf_det_points_fwd(None, None, None, None, None, None, 1, 1, 1, None, None) == -1
    one = ctypes.c_void_p(8)
    assert lib.adsurf_det_points_fwd(one, one, one, one, one, one, 0, 1, 1, one, None) == -2
    assert lib.adsurf_det_grid_bwd(one, one, one, one, one, one, None, 1, 500, 4, 4, None, one, one, one, one,
                                   one, 1 << 20, None) == -3
    assert lib.adsurf_det_grid_bwd(one, one, one, one, one, one, None, 1, 20, 4, 4, None, one, one, one, one,
                                   None, 0, None) == -4
    assert lib.adsurf_dmf_step(one, one, one, one, one, one, one, 1, 6, 4, 4, 9, 1, one, None, None, one, one, one, one,
                               one, 1 << 20, None) == -6


def test_product_has_no_cpu_path():
    import torch
    from adsurf_b200 import _capi
    x = torch.ones(1, 3, dtype=torch.float64)
    with pytest.raises(_capi.AdsurfError):
        _capi.points_fwd(x, x, x, x, x, x)


def test_product_does_not_import_oracle():
    """only tests/, smoke() and bench.py's CPU legs may touch oracle/"""
    pkg = os.path.join(ROOT, "adsurf_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), os.path.join(dirpath, f)
                assert "dltar_oracle" not in src, os.path.join(dirpath, f)
            if f.endswith((".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"#include\s+[<\"].*oracle", src), os.path.join(dirpath, f)
