"""Pin the oracle (oracle/dltar_oracle.py) against outputs of the unmodified reference.

The fixtures under tests/golden/ were produced by tests/golden/make_golden.py, which runs the
reference's own `_cps` evaluators, `multi_cal_grad`/`iter_cal_grad` and `surf96` in this
container.  CPU only.
"""
import json

import numpy as np
import pytest
import torch

from oracle import dltar_oracle as O
from tests.helpers import assert_det_close, assert_grad_close, load

TIGHT = 1e-12  # oracle vs reference, both ATen-CPU float64: only summation-order noise is expected


def T(x, grad=False):
    t = torch.tensor(np.asarray(x, dtype=np.float64))
    return t.requires_grad_(grad)


@pytest.mark.parametrize("name", ["grid_single_m6", "grid_single_m3", "grid_single_m20"])
def test_grid_single_fp64_and_fp32(name):
    g = load(name)
    det = O.det_grid(g["vlist"], g["tlist"], g["d"], g["a"], g["b"], g["rho"])
    assert det.shape == g["det64"].shape
    assert_det_close(det.numpy(), g["det64"], axis=0, rtol=TIGHT)
    # sign pattern of the determinant image is exact
    assert np.array_equal(np.sign(det.numpy()), np.sign(g["det64"]))
    # the reference's native float32 output is the float64 one to float32 accuracy (column-scaled)
    det32 = O.det_grid(g["vlist"], g["tlist"], g["d"], g["a"], g["b"], g["rho"], dtype=torch.float32)
    assert det32.dtype == torch.float32
    assert_det_close(det32.numpy(), g["det32"], axis=0, rtol=2e-3, what="float32 replay")


@pytest.mark.parametrize("name", ["points_m6", "points_m3", "points_m20", "points_m30"])
def test_points_batched_values_grads_jacobian(name):
    g = load(name)
    d, a, b, rho, c = (T(g[k], True) for k in ("d", "a", "b", "rho", "c"))
    F = O.det_points(c, T(g["t"]), d, a, b, rho)
    assert_det_close(F.detach().numpy(), g["F"], axis=-1, rtol=TIGHT, what="F")
    gc, gd, ga, gb, grho = torch.autograd.grad(F, (c, d, a, b, rho), grad_outputs=T(g["gF"]))
    for new, key in ((gd, "gd"), (ga, "ga"), (gb, "gb"), (grho, "grho"), (gc, "gc")):
        assert_grad_close(new.numpy(), g[key], rtol=1e-11, what=key)
    # single-model rank + per-sample Jacobians (sensitivity-kernel use, notebook 02_2 cell 12)
    d, a, b, rho, c = (T(g[k][0], True) for k in ("d", "a", "b", "rho", "c"))
    F1 = O.det_points(c, T(g["t"][0]), d, a, b, rho)
    assert F1.shape == (g["c"].shape[1],)
    eye = torch.eye(F1.numel(), dtype=torch.float64)
    for x, key in ((d, "Jd"), (a, "Ja"), (b, "Jb"), (rho, "Jrho"), (c, "Jc")):
        J = torch.autograd.grad(F1, x, grad_outputs=(eye,), retain_graph=True, is_grads_batched=True)[0]
        if key == "Jc":  # diagonal: dF_i/dc_i; scale = max over the picks of the model
            assert np.count_nonzero(J.numpy() - np.diag(np.diag(J.numpy()))) == 0
            assert_grad_close(np.diag(J.numpy()), np.diag(g[key]), rtol=1e-11, what=key)
        else:
            assert_grad_close(J.numpy(), g[key], rtol=1e-11, what=key)


@pytest.mark.parametrize("name", ["grid_m20", "grid_m6", "grid_m3"])
def test_grid_batched_values_and_grads(name):
    g = load(name)
    d, a, b, rho = (T(g[k], True) for k in ("d", "a", "b", "rho"))
    det = O.det_grid(g["vlist"], g["tlist"], d, a, b, rho)
    assert_det_close(det.detach().numpy(), g["det"], axis=1, rtol=TIGHT)
    assert np.array_equal(np.sign(det.detach().numpy()), np.sign(g["det"]))
    grads = torch.autograd.grad(det, (d, a, b, rho), grad_outputs=T(g["gdet"]), retain_graph=True)
    for new, key in zip(grads, ("gd", "ga", "gb", "grho")):
        assert_grad_close(new.numpy(), g[key], rtol=1e-11, what=key)
    grads = torch.autograd.grad(det, (d, a, b, rho), grad_outputs=torch.ones_like(det))
    for new, key in zip(grads, ("gd_ones", "ga_ones", "gb_ones", "grho_ones")):
        assert_grad_close(new.numpy(), g[key], rtol=1e-11, what=key)


@pytest.mark.parametrize("name", ["dmf_exp_m6", "dmf_log_m6", "dmf_raw_m20", "dmf_const_m3", "dmf_expnonorm_m6"])
def test_dmf_loss_and_grads(name):
    g = load(name)
    meta = json.loads(str(g["meta"]))
    b, d = T(g["b"], True), T(g["d"], True)
    loss = O.dmf_loss(g["c"], g["t"], g["C"], d, b, a=T(g["a"]), rho=T(g["rho"]),
                      initial_method=meta["initial_method"], vp_vs_ratio=meta["vp_vs_ratio"],
                      compress=meta["compress"], normalized=meta["normalized"],
                      compress_method=meta["compress_method"], damp_vertical=meta["damp_vertical"],
                      damp_horizontal=meta["damp_horizontal"])
    np.testing.assert_allclose(loss.detach().numpy(), g["loss"], rtol=1e-11)
    loss.backward(torch.ones_like(loss))
    assert_grad_close(b.grad.numpy(), g["gb"], rtol=1e-10, what="gb")
    if meta["inversion_method"] == "vs-and-thick":
        assert_grad_close(d.grad.numpy(), g["gd"], rtol=1e-10, what="gd")


def test_dmf_single_model():
    g = load("dmf_single_m6")
    b, d = T(g["b"][None], True), T(g["d"][None], True)
    loss = O.dmf_loss(g["c"][None], g["t"][None], g["C"][None], d, b, damp_vertical=0.002)
    np.testing.assert_allclose(loss.detach().numpy()[0], g["loss"], rtol=1e-11)
    loss.sum().backward()
    assert_grad_close(b.grad.numpy()[0], g["gb"], rtol=1e-10)
    assert_grad_close(d.grad.numpy()[0], g["gd"], rtol=1e-10)


def test_shipped_roots_are_zeros_of_the_oracle():
    """disper*.txt rows (reference's shipped known answers) are roots of the secular function.

    |F(T, c_root)| is tiny against the dense column range and the dense grid changes sign between
    the grid neighbours of every listed c.
    """
    s = load("shipped_fixtures")
    m6 = load("grid_single_m6")
    fr = s["friul"]
    from tests.golden_models import brocher_np
    cases = [
        (s["disper_00"][::25], m6["d"], m6["a"], m6["b"], m6["rho"], (2.0, 3.2)),
        (s["disper_01"][::40], *_model3(), (0.2, 0.4)),
        (s["disper_02"][::15], fr[:20, 1], brocher_np(fr[:20, 5])[0], fr[:20, 5], brocher_np(fr[:20, 5])[1], (0.5, 5.0)),
    ]
    for rows, d, a, b, rho, (vmin, vmax) in cases:
        t, c = rows[:, 0], rows[:, 1]
        F = O.det_points(c, t, d, a, b, rho).numpy()
        grid_c = np.linspace(vmin, vmax, 400)
        det = O.det_grid(grid_c, t, d, a, b, rho).numpy()  # [K, N]
        rng_ = det.max(0) - det.min(0)
        # float32-generated roots: c is accurate to ~3e-6 km/s, so |F|/range stays far below 1e-3
        assert np.max(np.abs(F) / rng_) < 1e-3
        eps = 2e-4 * (vmax - vmin)
        lo = O.det_points(c - eps, t, d, a, b, rho).numpy()
        hi = O.det_points(c + eps, t, d, a, b, rho).numpy()
        assert np.all(np.sign(lo) != np.sign(hi))


def _model3():
    from tests.golden_models import brocher_np
    b = np.array([0.3, 0.22, 0.4])
    a, rho = brocher_np(b)
    return np.array([0.005, 0.005, 0.005]), a, b, rho


def test_config2_first_loss_matches_shipped_trajectory():
    """model.json (float32 reference run) pins DMF loss at iteration 0 of config 2."""
    s = load("shipped_fixtures")
    dis = s["disper_00"]
    d, b = T(s["c2_init_thick"][None]), T(s["c2_init_vs"][None])
    C = np.arange(2, 3.2, 0.001)[None]
    loss = O.dmf_loss(dis[:, 1][None], dis[:, 0][None], C, d, b, damp_vertical=0.002)
    # float32 reference value 0.13665415; float64 oracle agrees to float32 accuracy
    assert abs(float(loss[0]) - float(s["c2_loss_head"][0])) < 2e-5
