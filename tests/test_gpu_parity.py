"""GPU parity tests: the CUDA path (through the C ABI) against the reference-generated golden
vectors and against the oracle on seeded inputs.  Tolerance: north_star's 1e-9 relative in FP64,
read as SURVEY.md §8(c) defines it (tests/helpers.py).  Sign patterns are compared exactly.
"""
import json

import numpy as np
import pytest
import torch

from tests.helpers import (COINCIDENT_RTOL, RTOL, assert_det_close, assert_grad_close, coincident_rows,
                           col_scaled_err, load)

pytestmark = pytest.mark.gpu

# two kernels compiled from the same source may contract FMAs differently: same value to rounding
SAME = 1e-12


@pytest.fixture(scope="module")
def ab():
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    import adsurf_b200
    from adsurf_b200 import _capi
    _capi.lib()  # must load: no fallback
    return adsurf_b200


def G(x):
    return torch.tensor(np.ascontiguousarray(x, dtype=np.float64), device="cuda")


def N(x):
    return x.detach().cpu().numpy()


# ---------------------------------------------------------------- golden: pointwise evaluator
@pytest.mark.parametrize("name", ["points_m6", "points_m3", "points_m20", "points_m30"])
def test_points_golden(ab, name):
    from adsurf_b200 import _capi
    g = load(name)
    d, a, b, rho, t, c, gF = (G(g[k]) for k in ("d", "a", "b", "rho", "t", "c", "gF"))
    F = _capi.points_fwd(d, a, b, rho, t, c)
    assert_det_close(N(F), g["F"], axis=-1, what="F")
    F2, gd, ga, gb, gr, gc = _capi.points_bwd(d, a, b, rho, t, c, gF, want_F=True)
    assert_det_close(N(F2), N(F), axis=-1, rtol=SAME, what="F (adjoint kernel)")
    for new, key in ((gd, "gd"), (ga, "ga"), (gb, "gb"), (gr, "grho"), (gc, "gc")):
        assert_grad_close(N(new), g[key], what=key)
    F3, Jd, Ja, Jb, Jr, Jc = _capi.points_jac(d, a, b, rho, t, c)
    assert_det_close(N(F3), N(F), axis=-1, rtol=SAME, what="F (jacobian kernel)")
    for J, key in ((Jd, "Jd"), (Ja, "Ja"), (Jb, "Jb"), (Jr, "Jrho")):
        assert_grad_close(N(J[0]), g[key], what=key)
    assert_grad_close(N(Jc[0]), np.diag(g["Jc"]), what="Jc")


@pytest.mark.parametrize("name", ["points_m6", "points_m20"])
def test_points_autograd_and_batched_grads(ab, name):
    """torch.autograd through the shim, incl. is_grads_batched (sensitivity notebook 02_2 cell 12)."""
    from adsurf_b200._cps import _surf96_vector_gpu as shim1
    from adsurf_b200._cps import _surf96_vectorAll_gpu as shimA
    g = load(name)
    d, a, b, rho, c = (G(g[k]).requires_grad_(True) for k in ("d", "a", "b", "rho", "c"))
    F = shimA.dltar_vector(c, G(g["t"]), d, a, b, rho, 2, -1, device="cuda")
    assert F.shape == g["F"].shape and F.dtype == torch.float64 and F.is_cuda
    gc, gd, ga, gb, gr = torch.autograd.grad(F, (c, d, a, b, rho), grad_outputs=G(g["gF"]))
    for new, key in ((gd, "gd"), (ga, "ga"), (gb, "gb"), (gr, "grho"), (gc, "gc")):
        assert_grad_close(N(new), g[key], what=key)
    # single-model rank + batched grad_outputs = eye
    d, a, b, rho, c = (G(g[k][0]).requires_grad_(True) for k in ("d", "a", "b", "rho", "c"))
    F1 = shim1.dltar_vector(c, G(g["t"][0]), d, a, b, rho, 2, -1, device="cuda")
    assert F1.shape == (g["c"].shape[1],)
    eye = torch.eye(F1.numel(), dtype=torch.float64, device="cuda")
    for x, key in ((d, "Jd"), (a, "Ja"), (b, "Jb"), (rho, "Jrho")):
        J = torch.autograd.grad(F1, x, grad_outputs=(eye,), retain_graph=True, is_grads_batched=True)[0]
        assert_grad_close(N(J), g[key], what=key)
    Jc = torch.autograd.grad(F1, c, grad_outputs=(eye,), retain_graph=True, is_grads_batched=True)[0]
    assert_grad_close(np.diag(N(Jc)), np.diag(g["Jc"]), what="Jc")
    # the result lands on the CPU when the caller asks for device="cpu" (computed on the GPU anyway)
    Fc = shim1.dltar_vector(g["c"][0], g["t"][0], g["d"][0], g["a"][0], g["b"][0], g["rho"][0], 2, -1)
    assert not Fc.is_cuda and torch.equal(Fc, F1.detach().cpu())  # same kernel, same inputs: deterministic


def test_sensitivity_kernel(ab):
    g = load("points_m30")
    out = ab.sensitivity(G(g["c"][0]), G(g["t"][0]), G(g["d"][0]), G(g["a"][0]), G(g["b"][0]), G(g["rho"][0]))
    ref = -g["Jb"] / np.diag(g["Jc"]).reshape(-1, 1)
    assert_grad_close(N(out["dc_dvs"]), ref, rtol=1e-8, what="dc/dvs")


# ---------------------------------------------------------------- golden: dense grid evaluator
@pytest.mark.parametrize("name", ["grid_m20", "grid_m6", "grid_m3"])
def test_grid_golden(ab, name):
    from adsurf_b200 import _capi
    g = load(name)
    d, a, b, rho, t, c, gdet = (G(g[k]) for k in ("d", "a", "b", "rho", "tlist", "vlist", "gdet"))
    det, cmax, cmin, bits = _capi.grid_fwd(d, a, b, rho, t, c, want_det=True, want_minmax=True, want_signbits=True)
    assert_det_close(N(det), g["det"], axis=1)
    assert np.array_equal(np.sign(N(det)), np.sign(g["det"]))  # sign-change / mode indices: exact
    assert torch.equal(cmax, det.max(dim=1).values) and torch.equal(cmin, det.min(dim=1).values)
    NF = t.shape[1]
    neg = N(det) < 0
    words = N(bits).astype(np.uint32)
    unpacked = ((words[..., None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(neg.shape[0], neg.shape[1], -1)[..., :NF]
    assert np.array_equal(unpacked.astype(bool), neg)
    det2, gd, ga, gb, gr = _capi.grid_bwd(d, a, b, rho, t, c, gdet, want_det=True)
    assert_det_close(N(det2), N(det), axis=1, rtol=SAME, what="det (adjoint kernel)")
    for new, key in zip((gd, ga, gb, gr), ("gd", "ga", "gb", "grho")):
        assert_grad_close(N(new), g[key], what=key)
    _, gd, ga, gb, gr = _capi.grid_bwd(d, a, b, rho, t, c, None)
    for new, key in zip((gd, ga, gb, gr), ("gd_ones", "ga_ones", "gb_ones", "grho_ones")):
        assert_grad_close(N(new), g[key], what=key)


@pytest.mark.parametrize("name", ["grid_single_m6", "grid_single_m3", "grid_single_m20"])
def test_grid_single_model_shim(ab, name):
    """[K,NF] rank through the dltar_matrix shim; grids that include c == layer velocity."""
    from adsurf_b200._cps import _surf96_matrix_gpu as shim
    g = load(name)
    det = shim.dltar_matrix(g["vlist"], g["tlist"], g["d"], g["a"], g["b"], g["rho"], 2, -1, device="Cuda")
    assert det.shape == g["det64"].shape and det.is_cuda
    det = N(det)
    co = coincident_rows(g["vlist"], g["a"], g["b"])
    scale = np.max(np.abs(g["det64"]), axis=0, keepdims=True)
    err = np.abs(det - g["det64"]) / scale
    assert err[~co].max() <= RTOL
    assert err[co].max() <= COINCIDENT_RTOL
    assert np.array_equal(np.sign(det), np.sign(g["det64"]))
    # the reference's own float32 output is reproduced to float32 accuracy
    assert col_scaled_err(det, g["det32"], axis=0) < 2e-3


def test_grid_autograd_through_shim(ab):
    from adsurf_b200._cps import _surf96_matrixAll_gpu as shim
    g = load("grid_m6")
    d, a, b, rho = (G(g[k]).requires_grad_(True) for k in ("d", "a", "b", "rho"))
    det = shim.dltar_matrix(G(g["vlist"]), G(g["tlist"]), d, a, b, rho, 2, -1, device="cuda:0")
    grads = torch.autograd.grad(det, (d, a, b, rho), grad_outputs=G(g["gdet"]))
    for new, key in zip(grads, ("gd", "ga", "gb", "grho")):
        assert_grad_close(N(new), g[key], what=key)


# ---------------------------------------------------------------- golden: Determinant Misfit Function
def _brocher(b):
    a = 0.9409 + 2.0947 * b - 0.8206 * b**2 + 0.2683 * b**3 - 0.0251 * b**4
    rho = 1.6612 * a - 0.4721 * a**2 + 0.0671 * a**3 - 0.0043 * a**4 + 0.000106 * a**5
    return a, rho


@pytest.mark.parametrize("name", ["dmf_exp_m6", "dmf_log_m6", "dmf_raw_m20", "dmf_const_m3", "dmf_expnonorm_m6"])
def test_dmf_golden(ab, name):
    from adsurf_b200._model import multi_cal_grad
    g = load(name)
    meta = json.loads(str(g["meta"]))
    calc = multi_cal_grad(G(g["d"]), G(g["a"]), G(g["b"]), G(g["rho"]), G(g["C"]), meta["damp_vertical"],
                          meta["damp_horizontal"], meta["compress"], meta["normalized"], meta["compress_method"],
                          meta["inversion_method"], meta["initial_method"], meta["vp_vs_ratio"], [], device="cuda")
    loss = calc(G(g["c"]), G(g["t"]))
    np.testing.assert_allclose(N(loss), g["loss"], rtol=1e-9)
    loss.backward(torch.ones_like(loss))
    assert_grad_close(N(calc.b.grad), g["gb"], what="gb")
    if meta["inversion_method"] == "vs-and-thick":
        assert_grad_close(N(calc.d.grad), g["gd"], what="gd")


def test_dmf_single_model_golden(ab):
    from adsurf_b200._model import iter_cal_grad
    g = load("dmf_single_m6")
    calc = iter_cal_grad(G(g["d"]), G(g["a"]), G(g["b"]), G(g["rho"]), G(g["C"]), 0.002, True, True, "exp",
                         "vs-and-thick", "Brocher", 2.45, [], device="cuda")
    loss = calc(G(g["c"]), G(g["t"]))
    np.testing.assert_allclose(float(loss), float(g["loss"]), rtol=1e-9)
    loss.backward()
    assert_grad_close(N(calc.b.grad), g["gb"], what="gb")
    assert_grad_close(N(calc.d.grad), g["gd"], what="gd")


# ---------------------------------------------------------------- oracle on seeded inputs (config-5 shape, reduced)
def _config5_subset(S, K, NF, seed=0):
    from tests.workload import config5_ensemble
    d, a, b, rho = config5_ensemble(S, seed=seed)
    t = np.tile(np.exp(np.linspace(np.log(1 / 20), np.log(5), NF)), (S, 1))
    c = np.tile(0.5 + 4.5 * (np.arange(K) * (1024 // K)) / 1024, (S, 1))
    return d, a, b, rho, t, c


def test_config5_subset_against_oracle(ab):
    """First 8 models x all 128 f x every 16th c of BASELINE config 5 through the float64 oracle."""
    from adsurf_b200 import _capi
    from oracle import dltar_oracle as O
    S, K, NF = 8, 64, 128
    d, a, b, rho, t, c = _config5_subset(S, K, NF)
    rng = np.random.default_rng(1)
    gdet = rng.standard_normal((S, K, NF))
    td, ta, tb, tr = (torch.tensor(x, requires_grad=True) for x in (d, a, b, rho))
    det_o = O.det_grid(torch.tensor(c), torch.tensor(t), td, ta, tb, tr)
    go = torch.autograd.grad(det_o, (td, ta, tb, tr), grad_outputs=torch.tensor(gdet), retain_graph=True)
    go1 = torch.autograd.grad(det_o, (td, ta, tb, tr), grad_outputs=torch.ones_like(det_o))
    det, gd, ga, gb, gr = _capi.grid_bwd(G(d), G(a), G(b), G(rho), G(t), G(c), G(gdet), want_det=True)
    assert_det_close(N(det), det_o.detach().numpy(), axis=1)
    assert np.array_equal(np.sign(N(det)), np.sign(det_o.detach().numpy()))
    for new, ref, key in zip((gd, ga, gb, gr), go, ("gd", "ga", "gb", "grho")):
        assert_grad_close(N(new), ref.numpy(), what=key)
    _, gd, ga, gb, gr = _capi.grid_bwd(G(d), G(a), G(b), G(rho), G(t), G(c), None)
    for new, ref, key in zip((gd, ga, gb, gr), go1, ("gd", "ga", "gb", "grho")):
        assert_grad_close(N(new), ref.numpy(), what=key + "(ones)")


@pytest.mark.parametrize("S,L,N_,K", [(1, 2, 1, 1), (3, 2, 33, 7), (2, 45, 70, 40), (5, 6, 500, 130), (1, 90, 31, 9),
                                      (40, 3, 174, 200), (17, 12, 65, 96), (150, 4, 40, 33), (2, 20, 128, 257),
                                      (9, 31, 97, 50), (64, 7, 20, 12), (1, 20, 300, 450), (300, 5, 31, 3)])
def test_ragged_shapes_against_oracle(ab, S, L, N_, K):
    """odd sizes: one sample, non-multiples of 32, a single finite layer, deep stacks (1 warp per block); shapes that
    land in the planner's other regimes (more models than SMs, shallow models, many / few trial velocities)."""
    from adsurf_b200 import _capi
    from oracle import dltar_oracle as O
    rng = np.random.default_rng(100 + L + 7 * S)
    b = np.sort(rng.uniform(1.0, 4.0, (S, L)), axis=1)
    d = rng.uniform(0.05, 0.6, (S, L))
    a, rho = _brocher(b)
    t = np.exp(rng.uniform(np.log(0.2), np.log(4), (S, N_)))
    c = rng.uniform(0.9, 4.2, (S, N_))
    C = np.sort(rng.uniform(0.9, 4.2, (S, K)), axis=1)
    gF = rng.standard_normal((S, N_))
    td, ta, tb, tr, tc = (torch.tensor(x, requires_grad=True) for x in (d, a, b, rho, c))
    Fo = O.det_points(tc, torch.tensor(t), td, ta, tb, tr)
    go = torch.autograd.grad(Fo, (td, ta, tb, tr, tc), grad_outputs=torch.tensor(gF))
    F, gd, ga, gb, gr, gc = _capi.points_bwd(G(d), G(a), G(b), G(rho), G(t), G(c), G(gF), want_F=True)
    assert_det_close(N(F), Fo.detach().numpy(), axis=-1)
    for new, ref, key in zip((gd, ga, gb, gr, gc), go, ("gd", "ga", "gb", "grho", "gc")):
        assert_grad_close(N(new), ref.numpy(), what=key)
    det_o = O.det_grid(C, t, d, a, b, rho).numpy()
    det, cmax, cmin, _ = _capi.grid_fwd(G(d), G(a), G(b), G(rho), G(t), G(C), want_minmax=True)
    assert_det_close(N(det), det_o, axis=1)
    assert np.array_equal(N(cmax), N(det).max(1)) and np.array_equal(N(cmin), N(det).min(1))
    _, gd2, ga2, gb2, gr2 = _capi.grid_bwd(G(d), G(a), G(b), G(rho), G(t), G(C), None)
    tdd, taa, tbb, trr = (torch.tensor(x, requires_grad=True) for x in (d, a, b, rho))
    O.det_grid(C, t, tdd, taa, tbb, trr).sum().backward()
    for new, ref, key in zip((gd2, ga2, gb2, gr2), (tdd, taa, tbb, trr), ("gd", "ga", "gb", "grho")):
        assert_grad_close(N(new), ref.grad.numpy(), what="grid " + key)


# ---------------------------------------------------------------- size-independent properties at larger sizes
def test_properties_at_scale(ab):
    """linearity in the upstream gradient, determinism, pointwise == dense on the same samples,
    fused max/min == max/min of the materialised grid; 64 models x 256 c x 128 f x 20 layers."""
    from adsurf_b200 import _capi
    S, K, NF = 64, 256, 128
    d, a, b, rho, t, c = (G(x) for x in _config5_subset(S, K, NF, seed=3))
    det, cmax, cmin, _ = _capi.grid_fwd(d, a, b, rho, t, c, want_minmax=True)
    assert torch.isfinite(det).all()
    assert torch.equal(cmax, det.max(1).values) and torch.equal(cmin, det.min(1).values)
    g1 = _capi.grid_bwd(d, a, b, rho, t, c, None)[1:]
    g1b = _capi.grid_bwd(d, a, b, rho, t, c, torch.ones_like(det))[1:]
    g2 = _capi.grid_bwd(d, a, b, rho, t, c, torch.full_like(det, 2.0))[1:]
    for x, y, z in zip(g1, g1b, g2):
        assert torch.equal(x, y)          # explicit ones == implicit ones, bit for bit (deterministic reduce)
        assert torch.equal(2.0 * x, z)    # exact linearity (power of two)
    # pointwise evaluator on the flattened grid of 4 models == dense evaluator
    s = slice(0, 4)
    cc = c[s].unsqueeze(2).expand(-1, -1, NF).reshape(4, -1).contiguous()
    tt = t[s].unsqueeze(1).expand(-1, K, -1).reshape(4, -1).contiguous()
    F, gd, ga, gb, gr, _ = _capi.points_bwd(d[s].contiguous(), a[s].contiguous(), b[s].contiguous(),
                                            rho[s].contiguous(), tt, cc, torch.ones_like(cc), want_F=True)
    assert_det_close(N(F.reshape(4, K, NF)), N(det[s]), axis=1, rtol=SAME)
    for x, y in zip((gd, ga, gb, gr), g1):
        assert_grad_close(N(x), N(y[s]), rtol=1e-11)


def test_nan_and_error_codes(ab):
    from adsurf_b200 import _capi
    g = load("points_m6")
    d, a, b, rho, t, c = (G(g[k]) for k in ("d", "a", "b", "rho", "t", "c"))
    bad = b.clone()
    bad[1, 2] = float("nan")
    F = _capi.points_fwd(d, a, bad, rho, t, c)
    assert torch.isnan(F[1]).all() and torch.isfinite(F[0]).all() and torch.isfinite(F[2]).all()
    lib = _capi.lib()
    assert lib.adsurf_det_points_fwd(None, None, None, None, None, None, 1, 1, 1, None, None) == -1
    assert lib.adsurf_workspace_bytes(_capi.OP_GRID_BWD, 4, 20, 128, 1024) > 0
    with pytest.raises(_capi.AdsurfError):
        _capi.points_fwd(d.cpu(), a.cpu(), b.cpu(), rho.cpu(), t.cpu(), c.cpu())
    L = 400  # adjoint state of 399 layers does not fit shared memory -> error code, not a crash
    big = torch.ones((1, L), dtype=torch.float64, device="cuda")
    with pytest.raises(_capi.AdsurfError):
        _capi.points_bwd(big, 2 * big, big, big, t[:1].contiguous(), c[:1].contiguous(), torch.ones_like(c[:1]).contiguous())
    F = _capi.points_fwd(big * 0.01, 2 * big, big, big, t[:1].contiguous(), c[:1].contiguous())  # forward handles it
    assert F.shape == (1, c.shape[1])


def test_periods_beyond_the_omega_clamp(ab):
    """omega is clamped at 1e-4 (periods beyond 2 pi 1e4 s).  The pointwise kernels follow the reference (k keeps the
    unclamped omega); the dense kernels evaluate at the clamped period, consistently in value and gradient."""
    from adsurf_b200 import _capi
    from oracle import dltar_oracle as O
    g = load("points_m6")
    d, a, b, rho = (g[k][:1] for k in ("d", "a", "b", "rho"))
    t = np.array([[1.0, 5.0e4, 7.0e4, 2.0e5, 1.0e6, 3.0]])
    c = np.array([[2.9, 3.1, 2.95, 3.3, 3.05, 3.2]])
    F = _capi.points_fwd(G(d), G(a), G(b), G(rho), G(t), G(c))
    Fo = O.det_points(torch.tensor(c), torch.tensor(t), *(torch.tensor(x) for x in (d, a, b, rho)))
    np.testing.assert_allclose(N(F), Fo.numpy(), rtol=1e-9)
    tc = np.minimum(t, 2.0 * np.pi / 1.0e-4)
    out_long = _capi.grid_bwd(G(d), G(a), G(b), G(rho), G(t), G(c), None, want_det=True)
    out_clamped = _capi.grid_bwd(G(d), G(a), G(b), G(rho), G(tc), G(c), None, want_det=True)
    for x, y in zip(out_long, out_clamped):
        np.testing.assert_allclose(N(x), N(y), rtol=1e-12, atol=0)
    det_o = O.det_grid(torch.tensor(c), torch.tensor(tc), *(torch.tensor(x) for x in (d, a, b, rho)))
    assert_det_close(N(out_clamped[0]), det_o.numpy(), axis=1)


def test_device_elementary_functions(ab):
    """inline rsqrt / exp(-x) / sincos of the kernels (device code paths) against libm: <= 2 ulp"""
    from adsurf_b200 import _capi
    rng = np.random.default_rng(5)
    x = np.concatenate([rng.uniform(0, 60, 200000), rng.uniform(0, 700, 20000), np.exp(rng.uniform(-30, 11.4, 200000)),
                        np.array([1e-300, 16.0, 59.999, 60.0, 99999.9, 1.0e5, 3.0e7])])
    rs, ex, sn, cs = (N(o) for o in _capi.math_probe(G(x)))
    ulp = lambda new, ref: np.abs(new - ref) / np.spacing(np.abs(ref))
    assert ulp(rs, 1 / np.sqrt(x)).max() <= 2
    m = x <= 690  # beyond ~700 the value is unspecified (callers use it only for p < 60)
    assert ulp(ex[m], np.exp(-x[m])).max() <= 2
    err = np.maximum(np.abs(sn - np.sin(x)), np.abs(cs - np.cos(x))) / np.spacing(1.0)
    assert err.max() <= 2


def test_full_size_config5_properties(ab):
    """BASELINE config 5 at its full size (4096 models x 128 f x 1024 c x 20 layers = 536 870 912 samples):
    size-independent properties of the fused forward+adjoint sweep.
      * batch independence: models 0..7 swept alone give the same det and gradients as inside the 4096 batch
        (and those 8 models are checked against the oracle by test_config5_subset_against_oracle);
      * exact linearity in the upstream gradient (gdet = 1 vs gdet = 2, power of two);
      * determinism (two runs bit-identical); fused max/min == max/min of the materialised grid;
      * row 0 of the ensemble is the unperturbed model: its column-wise sign-change count is >= 1 everywhere."""
    from adsurf_b200 import _capi
    from tests.workload import config5
    S, NF, K = 4096, 128, 1024
    d, a, b, rho, t, c = (G(x) for x in config5(S, NF, K))
    det, gd, ga, gb, gr = _capi.grid_bwd(d, a, b, rho, t, c, None, want_det=True)
    assert torch.isfinite(det).all() and all(torch.isfinite(x).all() for x in (gd, ga, gb, gr))
    again = _capi.grid_bwd(d, a, b, rho, t, c, None, want_det=False)[1:]
    for x, y in zip((gd, ga, gb, gr), again):
        assert torch.equal(x, y)
    sub = slice(0, 8)
    det8, gd8, ga8, gb8, gr8 = _capi.grid_bwd(*(x[sub].contiguous() for x in (d, a, b, rho, t, c)), None, want_det=True)
    assert torch.equal(det8, det[sub])
    for x, y in zip((gd8, ga8, gb8, gr8), (gd, ga, gb, gr)):
        assert_grad_close(N(x), N(y[sub]), rtol=1e-12)  # block partials are summed in a different order
    _, cmax, cmin, _ = _capi.grid_fwd(d, a, b, rho, t, c, want_det=False, want_minmax=True)
    assert_det_close(N(cmax[:64]), N(det[:64].max(1).values), axis=-1, rtol=SAME)
    assert_det_close(N(cmin[:64]), N(det[:64].min(1).values), axis=-1, rtol=SAME)
    sgn = torch.sign(det[0])
    assert int((sgn[:-1] != sgn[1:]).sum(0).min()) >= 1
    del det, det8
    two = torch.full((S, K, NF), 2.0, dtype=torch.float64, device="cuda")
    g2 = _capi.grid_bwd(d, a, b, rho, t, c, two)[1:]
    for x, y in zip((gd, ga, gb, gr), g2):
        assert torch.equal(2.0 * x, y)
