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


def test_grid_planner_covers_the_grid_and_fills_the_machine(lib):
    """host-side work split (no GPU: the planner then assumes 148 SMs)"""
    from adsurf_b200 import _capi
    for adj in (False, True):
        op = _capi.OP_GRID_BWD if adj else _capi.OP_GRID_FWD
        for S, L, NF, K in ((1, 6, 500, 1201), (100, 3, 174, 200), (4096, 20, 128, 1024), (512, 20, 128, 1024),
                            (1, 2, 1, 1), (3, 2, 33, 7), (2, 45, 70, 40), (1, 90, 31, 9), (7, 20, 300, 450)):
            p = _capi.plan_describe(op, S, L, NF, K)
            assert p["nft"] == (NF + 31) // 32 and 1 <= p["wpb"] <= (4 if adj else 8)
            assert p["wpb"] * p["nfg"] >= p["nft"] > p["wpb"] * (p["nfg"] - 1)
            assert p["nchunk"] * p["chunk"] >= K > (p["nchunk"] - 1) * p["chunk"]
            # table fills of up to 8 velocities (adjoint) / as many as the 26 KB table holds (forward, shallow models)
            assert 1 <= p["tc"] <= min(8 if adj else 64, p["chunk"]) and p["bpm"] == p["nchunk"] * p["nfg"]
            assert (p["tc"] * L + 1) * (20 if adj else 16) * 8 <= 26 * 1024 or p["tc"] == 1
    # config 2's dense pass (one model): about one wave of blocks instead of 24 serial velocities per thread
    p = _capi.plan_describe(_capi.OP_GRID_FWD, 1, 6, 500, 1201)
    assert p["chunk"] <= 12 and p["bpm"] * p["wpb"] <= 2 * p["sms"] * 16
    # the headline sweep keeps whole table fills and many waves
    p = _capi.plan_describe(_capi.OP_GRID_BWD, 4096, 20, 128, 1024)
    assert p["wpb"] == 4 and p["tc"] == 8 and p["chunk"] >= 64
    # config 3's dense pass (100 shallow models, several waves): blocks of two warps share one table fill
    p = _capi.plan_describe(_capi.OP_GRID_FWD, 100, 3, 174, 200)
    assert p["wpb"] == 2 and p["tc"] == p["chunk"] and 10 <= p["chunk"] <= 25


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
