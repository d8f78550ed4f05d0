"""Host-side logic on the CPU: autograd wrappers, the vmap rule behind is_grads_batched, the DMF
loss modules and the inversion loop bookkeeping — with the C-ABI calls served by the oracle
(tests/fake_backend.py).  These tests check the Python glue, not the kernels."""
import json
import types

import numpy as np
import pytest
import torch

from tests import fake_backend
from tests.helpers import assert_grad_close, load


@pytest.fixture()
def ab(monkeypatch):
    import adsurf_b200
    fake_backend.install(monkeypatch)
    return adsurf_b200


def T(x, grad=False):
    return torch.tensor(np.asarray(x, dtype=np.float64)).requires_grad_(grad)


def test_det_points_autograd_and_batched_grads(ab):
    g = load("points_m6")
    d, a, b, rho, c = (T(g[k], True) for k in ("d", "a", "b", "rho", "c"))
    F = ab.det_points(c, T(g["t"]), d, a, b, rho)
    np.testing.assert_allclose(F.detach().numpy(), g["F"], rtol=1e-12, atol=1e-12)
    gc, gd, ga, gb, gr = torch.autograd.grad(F, (c, d, a, b, rho), grad_outputs=T(g["gF"]))
    for new, key in ((gd, "gd"), (ga, "ga"), (gb, "gb"), (gr, "grho"), (gc, "gc")):
        assert_grad_close(new.numpy(), g[key], rtol=1e-11, what=key)
    # single-model rank + is_grads_batched goes through _PointsBwd.vmap -> Jacobian call
    d, a, b, rho, c = (T(g[k][0], True) for k in ("d", "a", "b", "rho", "c"))
    F1 = ab.det_points(c, T(g["t"][0]), d, a, b, rho)
    eye = torch.eye(F1.numel(), dtype=torch.float64)
    J = torch.autograd.grad(F1, b, grad_outputs=(eye,), retain_graph=True, is_grads_batched=True)[0]
    assert_grad_close(J.numpy(), g["Jb"], rtol=1e-11, what="Jb")
    Jc = torch.autograd.grad(F1, c, grad_outputs=(eye,), is_grads_batched=True)[0]
    assert_grad_close(np.diag(Jc.numpy()), np.diag(g["Jc"]), rtol=1e-11, what="Jc")


def test_det_grid_autograd_ranks(ab):
    g = load("grid_m6")
    d, a, b, rho = (T(g[k], True) for k in ("d", "a", "b", "rho"))
    det = ab.det_grid(g["vlist"], g["tlist"], d, a, b, rho)
    assert det.shape == g["det"].shape
    grads = torch.autograd.grad(det, (d, a, b, rho), grad_outputs=T(g["gdet"]))
    for new, key in zip(grads, ("gd", "ga", "gb", "grho")):
        assert_grad_close(new.numpy(), g[key], rtol=1e-11, what=key)
    g1 = load("grid_single_m6")
    det1 = ab.det_grid(g1["vlist"], g1["tlist"], g1["d"], g1["a"], g1["b"], g1["rho"])
    assert det1.shape == g1["det64"].shape  # [K, NF] for 1-D model arrays
    cmax, cmin = ab.det_grid_minmax(g1["vlist"], g1["tlist"], g1["d"], g1["a"], g1["b"], g1["rho"])
    assert cmax.shape == (g1["det64"].shape[1],)
    np.testing.assert_allclose(cmax.numpy(), g1["det64"].max(0), rtol=1e-12)


def test_shims_reject_unsupported_modes(ab):
    from adsurf_b200._cps import _surf96_vectorAll_gpu as v, _surf96_matrix as m
    g = load("points_m6")
    with pytest.raises(NotImplementedError):
        v.dltar_vector(g["c"], g["t"], g["d"], g["a"], g["b"], g["rho"], 1, -1)  # Love
    with pytest.raises(NotImplementedError):
        v.dltar_vector(g["c"], g["t"], g["d"], g["a"], g["b"], g["rho"], 3, -1)  # fast delta
    with pytest.raises(NotImplementedError):
        m.dltar_matrix(g["c"][0], g["t"][0], g["d"][0], g["a"][0], g["b"][0], g["rho"][0], 2, 0)  # water layer
    out = v.dltar_vector(g["c"], g["t"], g["d"], g["a"], g["b"], g["rho"], 2, -1, device="CPU")
    assert out.shape == g["F"].shape


@pytest.mark.parametrize("name", ["dmf_exp_m6", "dmf_log_m6", "dmf_raw_m20", "dmf_const_m3", "dmf_expnonorm_m6"])
def test_multi_cal_grad_module(ab, name):
    from adsurf_b200._model import multi_cal_grad
    g = load(name)
    meta = json.loads(str(g["meta"]))
    calc = multi_cal_grad(g["d"], g["a"], g["b"], g["rho"], g["C"], meta["damp_vertical"], meta["damp_horizontal"],
                          meta["compress"], meta["normalized"], meta["compress_method"], meta["inversion_method"],
                          meta["initial_method"], meta["vp_vs_ratio"], [], device="cpu")
    names = [n for n, _ in calc.named_parameters()]
    assert names == (["b", "d"] if meta["inversion_method"] == "vs-and-thick" else ["b"])  # optimiser order (_model.py:594-602)
    loss = calc(T(g["c"]), T(g["t"]))
    np.testing.assert_allclose(loss.detach().numpy(), g["loss"], rtol=1e-10)
    loss.backward(torch.ones_like(loss))
    assert_grad_close(calc.b.grad.numpy(), g["gb"], rtol=1e-9, what="gb")
    if meta["inversion_method"] == "vs-and-thick":
        assert_grad_close(calc.d.grad.numpy(), g["gd"], rtol=1e-9, what="gd")


def test_iter_cal_grad_module(ab):
    from adsurf_b200._model import iter_cal_grad
    g = load("dmf_single_m6")
    calc = iter_cal_grad(g["d"], g["a"], g["b"], g["rho"], g["C"], 0.002, True, True, "exp", "vs-and-thick", "Brocher",
                         2.45, [], device="cpu")
    loss = calc(g["c"], g["t"])
    assert loss.dim() == 0
    np.testing.assert_allclose(float(loss), float(g["loss"]), rtol=1e-10)
    loss.backward()
    assert_grad_close(calc.b.grad.numpy(), g["gb"], rtol=1e-9)
    assert_grad_close(calc.d.grad.numpy(), g["gd"], rtol=1e-9)


def _params(**kw):
    return types.SimpleNamespace(**kw)


def test_config2_trajectory_head(ab):
    """iter_inversion reproduces the head of the reference's shipped config-2 trajectory
    (examples/data/00_increase_layer/output/data/model.json: float32 run) to float32 accuracy:
    loss, Vs and thickness iterates through the first Adam steps and the depth clipping."""
    from adsurf_b200._model import iter_inversion
    s = load("shipped_fixtures")
    n_it = 4
    # every 10th pick and a 0.01 trial grid keep the oracle-backed run short; the full-size run is a GPU test
    obs = s["disper_00"]
    mp = _params(clist=np.arange(2, 3.2, 0.001), layering_method="LN", initialize_method="Brocher", vp_vs_ratio=2.45, rho=2)
    ip = _params(damp_vertical=0.002, damp_horizontal=0, compress=True, normalized=True, compress_method="exp",
                 inversion_method="vs-and-thick", lr=0.005, iteration=n_it, step_size=400, gamma=0.75, optimizer="Adam")
    wl = np.sort(obs[:, 0]) * np.sort(obs[:, 1])
    # Init_model's LN bounds (_ADsurf.py:455-470 -> depth_ln): wmin/3 and wmax/depth_factor for every layer
    from adsurf_b200._utils import depth_ln
    lo, hi = depth_ln(wl.min(), wl.max(), 6, 3.8)
    init = _params(init_model={"vs": s["c2_init_vs"], "vp": s["c2_init_vp"], "rho": s["c2_init_rho"],
                               "thick": s["c2_init_thick"], "layer_mindepth": lo, "layer_maxdepth": hi})
    inv = iter_inversion(mp, ip, init, obs, vsrange_sign="mul", vsrange=[0.1, 2], device="cpu")
    loss = np.asarray(inv.inv_process["loss"])
    np.testing.assert_allclose(loss[:n_it], s["c2_loss_head"][:n_it], atol=3e-5)
    np.testing.assert_allclose(np.asarray(inv.inv_process["iter_vs"])[:n_it], s["c2_iter_vs_head"][:n_it], atol=2e-5)
    np.testing.assert_allclose(np.asarray(inv.inv_process["iter_thick"])[:n_it], s["c2_iter_thick_head"][:n_it], atol=2e-5)
    assert set(inv.inv_model) == {"thick", "vp", "vs", "rho"}


def test_result_arrays_option(ab):
    """inv_param.result_arrays: the same numbers as the reference-style nested lists, as float64 arrays"""
    from adsurf_b200._model import iter_inversion, multi_inversion
    from tests.test_dist_gloo import _problem
    mp_, ip, init, obs = _problem()
    runs = {}
    for flag in (False, True):
        ip.result_arrays = flag
        inv = multi_inversion(mp_, ip, init, obs, vsrange_sign="add", vsrange=[-0.5, 0.5], device="cpu")
        runs[flag] = inv
    for key in ("iter_vs", "iter_thick", "loss"):
        lst, arr = runs[False].inv_process[key], runs[True].inv_process[key]
        assert isinstance(lst, list) and isinstance(arr, np.ndarray) and arr.dtype == np.float64
        assert np.array_equal(np.asarray(lst), arr)
    for key in ("thick", "vp", "vs", "rho"):
        assert isinstance(runs[False].inv_model[key], list)
        assert np.array_equal(np.asarray(runs[False].inv_model[key]), runs[True].inv_model[key])
    one = types.SimpleNamespace(init_model={k: np.asarray(v)[0] if np.ndim(v) == 2 else v
                                            for k, v in init.init_model.items()})
    ip.damp_horizontal = 0.0
    inv1 = iter_inversion(mp_, ip, one, obs[0], vsrange_sign="add", vsrange=[-0.5, 0.5], device="cpu")
    assert isinstance(inv1.inv_process["loss"], np.ndarray) and isinstance(inv1.inv_model["vs"], np.ndarray)


def test_gpu_dispersion_extraction_logic(ab):
    """dense-grid bracketing + batched regula falsi (adsurf_b200.dispersion) against the reference solver's
    roots; kernels served by the oracle here, the real kernels in tests/test_gpu_api.py."""
    from adsurf_b200.dispersion import dispersion_curves_dense as dispersion_curves
    g = load("surf96_roots")
    t, d, a, b, rho = (g[f"m3_{k}"] for k in ("t", "d", "a", "b", "rho"))
    sel = slice(0, 60, 4)
    c = dispersion_curves(t[sel], d, a, b, rho, modes=(0, 1, 2), dc=0.001)
    for i in range(3):
        ref = g[f"m3_mode{i}_c"][sel]
        both = (c[i] > 0) & (ref > 0)
        assert both.sum() >= 0.8 * (ref > 0).sum()
        assert np.max(np.abs(c[i][both] - ref[both])) < 5e-6  # the reference stops at 1e-6 relative


def test_bench_reference_arm_runs_on_the_cpu():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) prints one JSON line with the contract's
    keys; under torchrun only rank 0 works (other ranks exit 0 silently)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    rec = json.loads(lines[0])
    assert rec["impl"] == "reference" and rec["unit"] == "samples/s" and rec["value"] > 0 and rec["scaling"] == "strong"
    assert rec["cpu_baseline"]["kind"] == "port" and rec["cpu_baseline"]["cores"] >= 1
    assert rec["e2e"]["h2d_bytes_per_step"] == 0 and rec["gpu_launches"] == 0
    assert rec["metric"].startswith("DMF det samples/s fwd+bwd") and rec["dtype"] == "f64"
    quiet = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                            "--warmup", "0"], capture_output=True, text=True, env=dict(env, RANK="1", WORLD_SIZE="2"),
                           timeout=600)
    assert quiet.returncode == 0 and quiet.stdout.strip() == ""


def test_inversion_trajectories_match_the_reference_on_the_oracle(ab):
    """host logic + oracle against the reference's own float64 `inversion(...)` trajectories (configs 2-4 of
    tests/golden/trajectories.npz; config 3 on its first 8 models): what the GPU test checks with the real kernels"""
    import adsurf_b200._ADsurf as api
    from tests.trajectories import trajectory_errors
    errs = trajectory_errors(api, "cpu", c3_models=8)
    bad = {k: v for k, v in errs.items() if not v <= 1e-10}
    assert not bad, f"trajectories differ from the reference's: {bad} (all: {errs})"


def test_row_groups_slice_and_join_the_ensemble(ab):
    """_capi.InversionLoopGroup (host logic: contiguous row views of state / constants / picks per group, histories
    joined along the model axis, state updated in place) over the oracle-served loop: 3 groups of a 7-model ensemble
    give exactly the single loop's histories and final state; horizontal damping is refused."""
    from adsurf_b200 import _capi
    from tests.golden_models import brocher_np
    g = load("dmf_exp_m6")
    S, n_it = 7, 4
    rng = np.random.default_rng(9)
    b = np.tile(g["b"][0], (S, 1)) * (1 + rng.normal(0, 0.03, (S, 6)))
    a, rho = brocher_np(b)
    d = np.tile(g["d"][0], (S, 1))
    t, c = T(np.tile(g["t"][0][:10], (S, 1))), T(np.tile(g["c"][0][:10], (S, 1)))
    C = T(np.tile(np.arange(2, 3.2, 0.05), (S, 1)))
    kw = dict(compress=_capi.COMPRESS_EXP, vp_mode=1, vp_vs_ratio=0.0, invert_thick=True, layering=1,
              proj_flags=_capi.PROJ_LAST_LAYER, damp_vertical=0.002, damp_horizontal=0.0, lr=0.005, gamma=0.5,
              step_size=2, T=n_it)
    out = {}
    for groups in (1, 3):
        state = {"b": T(b), "d": T(d), "alpha": T(a), "rho": T(rho)}
        consts = {"b0": T(b), "alpha0": T(a), "rho0": T(rho), "lower_b": T(0.7 * b), "upper_b": T(1.3 * b),
                  "mindepth": T(np.cumsum(d, 1) * 0.8), "maxdepth": T(np.cumsum(d, 1) * 1.2)}
        loop = (_capi.InversionLoop(state, consts, t, c, C, **kw) if groups == 1 else
                _capi.InversionLoopGroup(state, consts, t, c, C, groups=groups, **kw))
        loop.run(n_it)
        assert loop.counter.tolist() == [n_it, 0]
        out[groups] = (loop.loss_hist, loop.iter_b_hist, loop.iter_d_hist, state["b"], state["d"], state["alpha"],
                       state["mb"], state["vd"])
        if groups == 3:
            assert [x.loss_hist.shape[1] for x in loop.loops] == [2, 2, 3] and tuple(loop.loss_hist.shape) == (n_it, S)
    for x, y in zip(out[1], out[3]):  # (the oracle's batched torch kernels round differently for other batch sizes:
        torch.testing.assert_close(x, y, rtol=1e-12, atol=1e-15)  # 1-ulp level; bit for bit on the GPU)
    with pytest.raises(_capi.AdsurfError):
        _capi.InversionLoopGroup(state, consts, t, c, C, groups=2, **dict(kw, damp_horizontal=0.001))
    assert _capi.loop_groups(4096, 20, 300, torch.device("cpu"), 0.0) == 1  # grouping is a GPU scheduling matter
