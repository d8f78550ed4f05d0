"""CPU check of the kernel arithmetic: adsurf_b200/csrc/adsurf_math.cuh compiled for the host.

The CUDA kernels and this test share one header, so the hand-derived adjoint (layer_adjoint,
half_adjoint, eig_adjoint) is verified against the reference-generated golden vectors here, without
a GPU.  The compiled helper is test infrastructure (tests/host_math/), not part of the product.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from tests.helpers import (COINCIDENT_RTOL, RTOL, assert_det_close, assert_grad_close, coincident_rows,
                           col_scaled_err, load)

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def hm(tmp_path_factory):
    out = tmp_path_factory.mktemp("host_math") / "libhost_math.so"
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wno-unknown-pragmas",
           "-I", os.path.join(ROOT, "adsurf_b200", "csrc"), os.path.join(HERE, "host_math", "host_math.cpp"),
           "-o", str(out)]
    subprocess.run(cmd, check=True)
    return ctypes.CDLL(str(out))


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def C(x):
    return np.ascontiguousarray(x, dtype=np.float64)


def run_points(hm, g, jac=False, seed_key="gF"):
    d, a, b, rho, t, c = (C(g[k]) for k in ("d", "a", "b", "rho", "t", "c"))
    S, L = d.shape
    N = c.shape[1]
    gF = C(g[seed_key]) if seed_key else None
    F = np.empty((S, N)); gc = np.empty((S, N))
    gd, ga, gb, gr = (np.empty((S, L)) for _ in range(4))
    J = [np.empty((S, N, L)) for _ in range(4)] if jac else [None] * 4
    rc = hm.host_points(P(d), P(a), P(b), P(rho), P(t), P(c), P(gF), S, L, N, P(F), P(gd), P(ga), P(gb), P(gr), P(gc),
                        *(P(j) for j in J))
    assert rc == 0
    return F, gd, ga, gb, gr, gc, J


@pytest.mark.parametrize("name", ["points_m6", "points_m3", "points_m20", "points_m30"])
def test_points_forward_adjoint_jacobian(hm, name):
    g = load(name)
    F, gd, ga, gb, gr, gc, J = run_points(hm, g, jac=True)
    assert_det_close(F, g["F"], axis=-1, what="F")
    for new, key in ((gd, "gd"), (ga, "ga"), (gb, "gb"), (gr, "grho"), (gc, "gc")):
        assert_grad_close(new, g[key], what=key)
    # Jacobian rows with unit seeds: re-run with seed 1
    F, gd, ga, gb, gr, gc, J = run_points(hm, g, jac=True, seed_key=None)
    for j, key in zip(J, ("Jd", "Ja", "Jb", "Jrho")):
        assert_grad_close(j[0], g[key], what=key)
    assert_grad_close(gc[0], np.diag(g["Jc"]), what="Jc")


@pytest.mark.parametrize("name", ["grid_m20", "grid_m6", "grid_m3"])
def test_grid_forward_adjoint(hm, name):
    g = load(name)
    d, a, b, rho, t, c, gdet = (C(g[k]) for k in ("d", "a", "b", "rho", "tlist", "vlist", "gdet"))
    S, L = d.shape
    K, NF = c.shape[1], t.shape[1]
    det = np.empty((S, K, NF))
    outs = [np.empty((S, L)) for _ in range(4)]
    assert hm.host_grid(P(d), P(a), P(b), P(rho), P(t), P(c), P(gdet), S, L, NF, K, P(det), *(P(o) for o in outs)) == 0
    assert_det_close(det, g["det"], axis=1)
    assert np.array_equal(np.sign(det), np.sign(g["det"]))
    for new, key in zip(outs, ("gd", "ga", "gb", "grho")):
        assert_grad_close(new, g[key], what=key)
    assert hm.host_grid(P(d), P(a), P(b), P(rho), P(t), P(c), None, S, L, NF, K, P(det), *(P(o) for o in outs)) == 0
    for new, key in zip(outs, ("gd_ones", "ga_ones", "gb_ones", "grho_ones")):
        assert_grad_close(new, g[key], what=key)


@pytest.mark.parametrize("name", ["grid_single_m6", "grid_single_m3", "grid_single_m20"])
def test_grid_single_sign_pattern(hm, name):
    """includes trial velocities that coincide with layer velocities (c == beta_m) in the forward"""
    g = load(name)
    d, a, b, rho = (C(g[k][None]) for k in ("d", "a", "b", "rho"))
    t, c = C(g["tlist"][None]), C(g["vlist"][None])
    K, NF, L = c.shape[1], t.shape[1], d.shape[1]
    det = np.empty((1, K, NF))
    outs = [np.empty((1, L)) for _ in range(4)]
    assert hm.host_grid(P(d), P(a), P(b), P(rho), P(t), P(c), None, 1, L, NF, K, P(det), *(P(o) for o in outs)) == 0
    co = coincident_rows(g["vlist"], g["a"], g["b"])
    assert co.any()
    scale = np.max(np.abs(g["det64"]), axis=0, keepdims=True)
    err = np.abs(det[0] - g["det64"]) / scale
    assert err[~co].max() <= RTOL
    assert err[co].max() <= COINCIDENT_RTOL
    assert np.array_equal(np.sign(det[0]), np.sign(g["det64"]))


def _ulps(new, ref):
    ref = np.asarray(ref, dtype=np.float64)
    return np.abs(new - ref) / np.spacing(np.abs(ref))


def test_inline_elementary_functions_accuracy(hm):
    """exp_neg / sincos_fast of adsurf_math.cuh against numpy's libm: a few ulp at most."""
    rng = np.random.default_rng(5)
    x = np.concatenate([rng.uniform(0, 60, 20000), rng.uniform(0, 700, 5000), np.exp(rng.uniform(-30, 11.4, 20000)),
                        np.array([0.0, 1e-300, 16.0, 59.999, 60.0, 99999.9, 1.0e5, 3.0e7])])
    n = len(x)
    rs, ex, sn, cs = (np.empty(n) for _ in range(4))
    hm.host_elementary(P(C(x)), n, P(rs), P(ex), P(sn), P(cs))
    pos = x > 0
    assert _ulps(rs[pos], 1 / np.sqrt(x[pos])).max() <= 2
    m = x <= 690  # beyond ~700 the value is unspecified (callers use it only for p < 60)
    assert _ulps(ex[m], np.exp(-x[m])).max() <= 2
    # sin/cos: error measured against the larger of the two (the reduction error is absolute)
    err = np.maximum(np.abs(sn - np.sin(x)), np.abs(cs - np.cos(x))) / np.spacing(1.0)
    assert err.max() <= 2


@pytest.mark.parametrize("seed", range(12))
def test_random_models_against_oracle(hm, seed):
    """Seeded random layered models (2-14 layers, low-velocity zones allowed, trial velocities in every regime:
    below all Vs, between Vs and Vp, above Vp of the top layers) — kernel arithmetic (host build) vs the oracle's
    autograd, dense (table) path and pointwise (direct) path."""
    import torch
    from oracle import dltar_oracle as O
    rng = np.random.default_rng(1000 + seed)
    S, L = 2, int(rng.integers(2, 15))
    b = rng.uniform(0.3, 4.5, (S, L))
    a = b * rng.uniform(1.5, 2.2, (S, L))
    rho = rng.uniform(1.5, 3.3, (S, L))
    d = rng.uniform(0.02, 0.8, (S, L))  # phases up to ~250 rad; beyond that 1-ulp input noise alone exceeds 1e-9
    K, NF = 24, 7
    c = np.sort(rng.uniform(0.25, 8.0, (S, K)), axis=1)
    t = np.exp(rng.uniform(np.log(0.1), np.log(8.0), (S, NF)))
    gdet = rng.standard_normal((S, K, NF))
    td, ta, tb, tr = (torch.tensor(x, requires_grad=True) for x in (d, a, b, rho))
    det_o = O.det_grid(torch.tensor(c), torch.tensor(t), td, ta, tb, tr)
    go = torch.autograd.grad(det_o, (td, ta, tb, tr), grad_outputs=torch.tensor(gdet))
    det = np.empty((S, K, NF))
    outs = [np.empty((S, L)) for _ in range(4)]
    assert hm.host_grid(P(C(d)), P(C(a)), P(C(b)), P(C(rho)), P(C(t)), P(C(c)), P(C(gdet)), S, L, NF, K, P(det),
                        *(P(o) for o in outs)) == 0
    assert_det_close(det, det_o.detach().numpy(), axis=1)
    assert np.array_equal(np.sign(det), np.sign(det_o.detach().numpy()))
    for new, ref, key in zip(outs, go, ("gd", "ga", "gb", "grho")):
        assert_grad_close(new, ref.numpy(), what=key)
    # pointwise path on the flattened grid of model 0, including dF/dc
    cc = np.repeat(c[:1], NF, axis=0).T.reshape(1, -1)
    tt = np.tile(t[:1], (K, 1)).reshape(1, -1)
    tc = torch.tensor(cc, requires_grad=True)
    td, ta, tb, tr = (torch.tensor(x[:1], requires_grad=True) for x in (d, a, b, rho))
    Fo = O.det_points(tc, torch.tensor(tt), td, ta, tb, tr)
    gF = rng.standard_normal(cc.shape)
    go = torch.autograd.grad(Fo, (td, ta, tb, tr, tc), grad_outputs=torch.tensor(gF))
    # conditioning of this random case: what a 1-ulp change of the thicknesses alone does to the oracle's own
    # gradients.  Where that exceeds the tolerance the comparison is bounded by it (x4) instead of by 1e-9.
    td1 = torch.tensor(np.nextafter(d[:1], np.inf), requires_grad=True)
    Fn = O.det_points(tc, torch.tensor(tt), td1, ta, tb, tr)
    gn = torch.autograd.grad(Fn, (td1, ta, tb, tr, tc), grad_outputs=torch.tensor(gF))
    shp = lambda x, i: x.numpy().reshape(K, NF).T if i == 4 else x.numpy()
    tol = [max(RTOL, 4.0 * col_scaled_err(shp(x, i), shp(y, i), axis=-1)) for i, (x, y) in enumerate(zip(gn, go))]
    g = {"d": d[:1], "a": a[:1], "b": b[:1], "rho": rho[:1], "t": tt, "c": cc, "gF": gF}
    F, gd, ga, gb, gr, gc, _ = run_points(hm, g)
    assert_det_close(F.reshape(K, NF), Fo.detach().numpy().reshape(K, NF), axis=0)
    for i, (new, ref, key) in enumerate(zip((gd, ga, gb, gr), go[:4], ("gd", "ga", "gb", "grho"))):
        assert_grad_close(new, ref.numpy(), rtol=tol[i], what="points " + key)
    assert_grad_close(gc.reshape(K, NF).T, go[4].numpy().reshape(K, NF).T, rtol=tol[4], what="gc")


def test_pointwise_period_beyond_omega_clamp(hm):
    """periods beyond 2 pi 1e4 s: the reference clamps omega to 1e-4 but not k (_surf96_vectorAll_gpu.py:319,326),
    so k/omega depends on the period; the pointwise path follows that exactly (value and gradients)."""
    import torch
    from oracle import dltar_oracle as O
    g = load("points_m6")
    d, a, b, rho = (C(g[k][:1]) for k in ("d", "a", "b", "rho"))
    t = C(np.array([[1.0, 5.0e4, 7.0e4, 2.0e5, 1.0e6, 3.0]]))
    c = C(np.array([[2.9, 3.1, 2.95, 3.3, 3.05, 3.2]]))
    gF = C(np.array([[1.0, -2.0, 0.5, 1.5, -1.0, 2.0]]))
    td, ta, tb, tr, tc = (torch.tensor(x, requires_grad=True) for x in (d, a, b, rho, c))
    Fo = O.det_points(tc, torch.tensor(t), td, ta, tb, tr)
    go = torch.autograd.grad(Fo, (td, ta, tb, tr, tc), grad_outputs=torch.tensor(gF))
    F, gd, ga, gb, gr, gc, _ = run_points(hm, {"d": d, "a": a, "b": b, "rho": rho, "t": t, "c": c, "gF": gF})
    np.testing.assert_allclose(F, Fo.detach().numpy(), rtol=1e-9)
    for new, ref, key in zip((gd, ga, gb, gr, gc), go, ("gd", "ga", "gb", "grho", "gc")):
        assert_grad_close(new, ref.numpy(), what=key)


@pytest.mark.parametrize("ifunc", [2, 1])
@pytest.mark.parametrize("tag,nmodes", [("m6", 1), ("m3", 3), ("m20", 1), ("w5", 2), ("w5s", 1)])
def test_dispersion_search_matches_reference_solver(hm, tag, nmodes, ifunc):
    """adsurf_dispersion.cuh (what k_dispersion runs per thread) compiled for the host: the reference solver's roots
    (tests/golden/surf96_roots.npz) within 1e-6 km/s, identical zeros for absent modes — and the same numbers as the
    Python host port `_surf96.surf96` to rounding (same control flow).  ifunc 2: Rayleigh (dltar4), 1: Love (dltar1); w5 / w5s: a water layer on top."""
    from adsurf_b200._surf96 import surf96
    g = load("surf96_roots")
    t, d, a, b, rho = (C(g[f"{tag}_{k}"]) for k in ("t", "d", "a", "b", "rho"))
    dc = float(g[f"{tag}_dc"])
    NF, L = len(t), len(d)
    c = np.empty((nmodes, NF))
    status = np.zeros(1, dtype=np.int32)
    hm.host_dispersion.argtypes = ([ctypes.c_void_p] * 5 + [ctypes.c_int] * 4 +
                                   [ctypes.c_double, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p])
    hm.host_dispersion(P(d), P(a), P(b), P(rho), P(t), 1, L, NF, nmodes, dc, ifunc, P(c),
                       status.ctypes.data_as(ctypes.c_void_p))
    assert status[0] == 0
    for mode in range(nmodes):
        ref = g[f"{tag}_mode{mode}_c" if ifunc == 2 else f"{tag}_love_mode{mode}_c"]
        assert np.array_equal(c[mode] > 0, ref > 0)
        assert np.max(np.abs(c[mode] - ref)) <= 1e-6
        port = surf96(t, d, a, b, rho, mode, 0, ifunc, dc)
        assert np.max(np.abs(c[mode] - port)) <= 1e-9
