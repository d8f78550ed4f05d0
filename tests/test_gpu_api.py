"""GPU tests of the reference-facing API (BASELINE configs 1-4 at their real sizes, short runs):
`inversion(...)`, `init_model_MonteCarlo`, `cal_determinant`, through adsurf_b200._ADsurf."""
import numpy as np
import pytest
import torch

from tests.helpers import assert_det_close, load

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    import adsurf_b200._ADsurf as A
    return A


def _mp6(A):
    return A.Model_param(dc=0.001, vmin=2, vmax=3.2, tmin=1 / 10, tmax=5, layering_method="LN",
                         initialize_method="Brocher", layering_ratio=1.2, depth_factor=3.8, layer_number=6,
                         vp_vs_ratio=2.45, rho=2)


def test_config1_determinant_image(api):
    """cal_determinant on the 6-layer model: [1201 x 500] image, sign pattern vs the oracle on a sub-grid"""
    from oracle import dltar_oracle as O
    g = load("grid_single_m6")
    true_model = api.True_model(_mp6(api), thick=g["d"], vs=g["b"], vp=g["a"], rho=g["rho"])
    F = true_model._cal_determinant(sampling_method="log-wavelength", sampling_num=500)
    assert tuple(F.shape) == (1200, 500) or tuple(F.shape) == (1201, 500)
    vl, tl = np.asarray(true_model.vlist), np.asarray(true_model.tlist)
    ks, fs = np.arange(5, len(vl), 40), np.arange(0, 500, 25)   # off the layer velocities (k % 10 == 5)
    ref = O.det_grid(vl[ks], tl[fs], g["d"], g["a"], g["b"], g["rho"]).numpy()
    sub = F.numpy()[np.ix_(ks, fs)]
    assert_det_close(sub, ref, axis=0)
    assert np.array_equal(np.sign(sub), np.sign(ref))


def test_config2_single_model_inversion_head(api):
    """full-size config 2 (N=500 picks, K=1200 trial c, vs-and-thick, LN clipping): first iterations
    against the reference's shipped float32 trajectory (examples/data/00_increase_layer/output/data/model.json)"""
    s = load("shipped_fixtures")
    obs = s["disper_00"]
    mp = _mp6(api)
    ip = api.inv_param(inversion_method="vs-and-thick", compress=True, compress_method="exp", normalized=True, lr=0.005,
                       damp_vertical=0.002, damp_horizontal=0, iteration=8, step_size=400, gamma=0.75)
    init = api.Init_model(model_param=mp, pvs_obs=obs, thick=[0.4, 0.5, 0.5, 0.5, 0.5, 0.5])
    np.testing.assert_allclose(init.init_model["vs"], s["c2_init_vs"], rtol=1e-6)
    np.testing.assert_allclose(init.init_model["thick"], s["c2_init_thick"], rtol=1e-6)
    inv = api.inversion(model_param=mp, inv_param=ip, init_model=init, pvs_obs=obs, vsrange_sign="mul",
                        vsrange=[0.1, 2], AK135_data=[], device="Cuda")
    np.testing.assert_allclose(inv.inv_process["loss"][:8], s["c2_loss_head"][:8], atol=3e-5)
    np.testing.assert_allclose(np.asarray(inv.inv_process["iter_vs"])[:8], s["c2_iter_vs_head"][:8], atol=3e-5)
    np.testing.assert_allclose(np.asarray(inv.inv_process["iter_thick"])[:8], s["c2_iter_thick_head"][:8], atol=3e-5)


def test_config3_ensemble_inversion_and_prescreen(api):
    """100-model multimodal sandwich inversion (3 layers, 174 unlabelled picks, 200 trial c): Monte-Carlo
    pre-screen and the first DMF losses against the oracle; loss decreases over a short run."""
    from oracle import dltar_oracle as O
    s = load("shipped_fixtures")
    obs = s["disper_01_obs"]
    mp = api.Model_param(dc=0.001, vmin=0.2, vmax=0.4, tmin=1 / 100, tmax=1 / 5, layering_method="LN",
                         initialize_method="Constant", layering_ratio=2.5, depth_factor=2.5, layer_number=3,
                         vp_vs_ratio=2.45, rho=2, fundamental_range=[1 / 37, 1 / 8])
    ip = api.inv_param(inversion_method="vs-and-thick", compress=True, compress_method="exp", normalized=True, lr=0.001,
                       damp_vertical=0.0, damp_horizontal=0.0, iteration=12, step_size=200, gamma=0.75)
    init = api.Init_model(model_param=mp, pvs_obs=obs, thick=np.array([0.005, 0.005, 0.005]))
    np.random.seed(123)
    mc = api.init_model_MonteCarlo(mp, init, sampling_num=100, sampling_method="normal", vsrange_sign="plus",
                                   vsrange=[-0.1, 0.2], sigma_vs=20, sigma_thick=20)
    M = mc.MonteCarlo_model
    assert M["vs"].shape == (100, 3) and len(M["loss"]) == 100
    # pre-screen value of three models against the oracle: mean |F / (max - min)|
    for i in (0, 7, 99):
        F = O.det_points(obs[:, 1], obs[:, 0], M["thick"][i], M["vp"][i], M["vs"][i], M["rho"][i])
        det = O.det_grid(mp.clist, obs[:, 0], M["thick"][i], M["vp"][i], M["vs"][i], M["rho"][i])
        ref = float(torch.mean(torch.abs(F / (det.max(0).values - det.min(0).values))))
        assert abs(M["loss"][i] - ref) <= 1e-9 * max(ref, 1.0)
    for key in ("vs", "vp", "thick", "rho"):
        init.init_model[key] = M[key]
    S = 100
    pvs = np.ones((S, obs.shape[0], 2)) * obs[:, :2]
    inv = api.inversion(model_param=mp, inv_param=ip, init_model=init, pvs_obs=pvs, vsrange_sign="plus",
                        vsrange=[-0.8, 0.8], AK135_data=[], device="cuda")
    loss = np.asarray(inv.inv_process["loss"])
    assert loss.shape == (12, S) and np.isfinite(loss).all()
    ref0 = O.dmf_loss(pvs[:3, :, 1], pvs[:3, :, 0], np.tile(mp.clist, (3, 1)), torch.tensor(M["thick"][:3]),
                      torch.tensor(M["vs"][:3]), a=torch.tensor(M["vp"][:3]), rho=torch.tensor(M["rho"][:3]),
                      initial_method="Constant", vp_vs_ratio=2.45).numpy()
    np.testing.assert_allclose(loss[0, :3], ref0, rtol=1e-9)
    assert np.median(loss[-1]) < np.median(loss[0])
    assert np.asarray(inv.inv_model["vs"]).shape == (S, 3)


def test_config4_friul_ensemble_step_and_sensitivity(api):
    """FRIUL7W 20-layer ensemble of 1024 models, N=300 picks, compress=False: one optimiser step; plus the
    30-layer sensitivity kernels at the 300 shipped picks (notebook 02_2)."""
    from oracle import dltar_oracle as O
    from adsurf_b200 import ops
    from adsurf_b200.workload import config5_ensemble, friul7w
    s = load("shipped_fixtures")
    obs = s["disper_02"]
    d, a, b, rho = config5_ensemble(1024, 20, seed=0)
    mp = api.Model_param(dc=0.01, vmin=0.5, vmax=5.0, tmin=1 / 20, tmax=5, layering_method="None",
                         initialize_method="Brocher", layering_ratio=1.2, depth_factor=2, layer_number=6,
                         vp_vs_ratio=2.45, rho=2)
    ip = api.inv_param(inversion_method="vs", compress=False, compress_method="exp", normalized=True, lr=0.03,
                       damp_vertical=0.0, damp_horizontal=0.0, iteration=3, step_size=200, gamma=0.75)
    init = api.Init_model(model_param=mp, pvs_obs=obs, vp=a[0], vs=b[0], rho=rho[0], thick=d[0])
    for key, val in (("vs", b), ("vp", a), ("thick", d), ("rho", rho)):
        init.init_model[key] = val
    pvs = np.ones((1024, obs.shape[0], 2)) * obs[:, :2]
    inv = api.inversion(model_param=mp, inv_param=ip, init_model=init, pvs_obs=pvs, vsrange_sign="mul",
                        vsrange=[0.1, 2.5], device="cuda:0")
    loss = np.asarray(inv.inv_process["loss"])
    assert loss.shape == (3, 1024) and np.isfinite(loss).all()
    ref = O.dmf_loss(pvs[:2, :, 1], pvs[:2, :, 0], None, torch.tensor(d[:2]), torch.tensor(b[:2]), compress=False).numpy()
    np.testing.assert_allclose(loss[0, :2], ref, rtol=1e-9)
    # sensitivity kernels dc/dVs on the 30-layer model
    th, vp, vs, rr = friul7w(30)
    out = ops.sensitivity(obs[:, 1], obs[:, 0], th, vp, vs, rr)
    assert tuple(out["dc_dvs"].shape) == (300, 30) and torch.isfinite(out["dc_dvs"]).all()
    tv, tc = torch.tensor(vs, requires_grad=True), torch.tensor(obs[:40, 1].copy(), requires_grad=True)
    F = O.det_points(tc, obs[:40, 0], th, vp, tv, rr)
    eye = torch.eye(40, dtype=torch.float64)
    Jb = torch.autograd.grad(F, tv, grad_outputs=(eye,), retain_graph=True, is_grads_batched=True)[0]
    Jc = torch.diagonal(torch.autograd.grad(F, tc, grad_outputs=(eye,), is_grads_batched=True)[0])
    kern = (-Jb / Jc[:, None]).numpy()
    got = out["dc_dvs"][:40].cpu().numpy()
    scale = np.max(np.abs(kern), axis=1, keepdims=True)
    assert np.max(np.abs(got - kern) / scale) < 1e-6   # roots are float32-accurate: dF/dc is well away from 0


@pytest.mark.parametrize("tag,modes", [("m6", (0,)), ("m3", (0, 1, 2)), ("m20", (0,)), ("w5", (0, 1)), ("w5s", (0,))])
def test_gpu_dispersion_curves_match_reference_solver(api, tag, modes):
    """k_dispersion (the reference's root search with its mode bookkeeping, one warp per model) against the
    reference's sequential solver in float64: every period, every mode, 1e-6 km/s; identical zeros for absent modes."""
    from adsurf_b200.dispersion import dispersion_curves
    g = load("surf96_roots")
    t, d, a, b, rho = (g[f"{tag}_{k}"] for k in ("t", "d", "a", "b", "rho"))
    c = dispersion_curves(t, d, a, b, rho, modes=modes, dc=float(g[f"{tag}_dc"]))
    assert c.shape == (len(modes), len(t))
    for i, mode in enumerate(modes):
        ref = g[f"{tag}_mode{mode}_c"]
        assert np.array_equal(c[i] > 0, ref > 0)           # mode indices / absent modes: exact
        assert np.max(np.abs(c[i] - ref)) <= 1e-6          # km/s, 100 % of the points


@pytest.mark.parametrize("tag,modes", [("m6", (0,)), ("m3", (0, 1, 2)), ("m20", (0,)), ("w5", (0, 1)), ("w5s", (0,))])
def test_gpu_love_dispersion_curves_match_reference_solver(api, tag, modes):
    """ifunc = 1: Love waves (the reference's dltar1 behind the same root search) on the GPU"""
    from adsurf_b200.dispersion import dispersion_curves
    g = load("surf96_roots")
    t, d, a, b, rho = (g[f"{tag}_{k}"] for k in ("t", "d", "a", "b", "rho"))
    c = dispersion_curves(t, d, a, b, rho, modes=modes, dc=float(g[f"{tag}_dc"]), ifunc=1)
    for i, mode in enumerate(modes):
        ref = g[f"{tag}_love_mode{mode}_c"]
        assert np.array_equal(c[i] > 0, ref > 0)
        assert np.max(np.abs(c[i] - ref)) <= 1e-6
    with pytest.raises(NotImplementedError):
        dispersion_curves(t, d, a, b, rho, modes=modes, dc=0.01, ifunc=3)


def test_gpu_group_velocity_matches_host_solver(api):
    """itype=1 (group velocity from two phase curves, _surf96.py:735-758) on the GPU against the host solver"""
    from adsurf_b200._surf96 import surf96
    from adsurf_b200.dispersion import dispersion_curves
    g = load("surf96_roots")
    t, d, a, b, rho = (g[f"m6_{k}"] for k in ("t", "d", "a", "b", "rho"))
    u = dispersion_curves(t[::6], d, a, b, rho, modes=(0,), dc=0.005, itype=1)
    ref = surf96(t[::6], d, a, b, rho, 0, 1, 2, 0.005)
    assert u.shape == (1, len(t[::6])) and np.all(u > 0)
    np.testing.assert_allclose(u[0], ref, rtol=0, atol=1e-6)


def test_gpu_dispersion_ensemble_equals_single_models(api):
    """the parallel axis is the ensemble: 96 perturbed 6-layer models in one launch == 96 single-model launches"""
    from adsurf_b200.dispersion import dispersion_curves
    from tests.golden_models import brocher_np
    g = load("surf96_roots")
    rng = np.random.default_rng(4)
    S = 96
    b = g["m6_b"] * (1 + rng.normal(0, 0.02, (S, 6)))
    a, rho = brocher_np(b)
    d = np.tile(g["m6_d"], (S, 1))
    c = dispersion_curves(g["m6_t"], d, a, b, rho, modes=(0, 1), dc=0.005)
    assert c.shape == (S, 2, len(g["m6_t"])) and np.all(c[:, 0] > 0)
    for s in (0, 17, 95):
        one = dispersion_curves(g["m6_t"], d[s], a[s], b[s], rho[s], modes=(0, 1), dc=0.005)
        assert np.array_equal(one, c[s])


def test_gpu_dispersion_more_layers_than_lanes(api):
    """45 layers: the lanes of the model's warp stride over the layers of an evaluation (k_dispersion / DispWarpEval);
    same roots as the sequential host solver"""
    from adsurf_b200._surf96 import surf96
    from adsurf_b200._utils import gen_model
    from adsurf_b200.dispersion import dispersion_curves
    L = 45
    vs = np.linspace(1.0, 4.2, L) + 0.15 * np.sin(np.arange(L))
    th, vp, vs, rho = gen_model(np.full(L, 0.35), vs, area=True)
    t = np.exp(np.linspace(np.log(0.2), np.log(12.0), 24))
    c = dispersion_curves(t, th, vp, vs, rho, modes=(0, 1), dc=0.005)
    for mode in (0, 1):
        ref = surf96(t, th, vp, vs, rho, mode, 0, 2, 0.005)
        assert np.array_equal(c[mode] > 0, ref > 0)
        assert np.max(np.abs(c[mode] - ref)) <= 1e-8


@pytest.mark.parametrize("name,tol", [("disper_00", 3e-6), ("disper_01", 3e-6), ("disper_02", 6e-6)])
def test_cal_dispersion_curve_on_gpu_reproduces_shipped_fixtures(api, name, tol):
    """`cal_dispersion_curve(..., device="cuda")` against the reference's shipped disper*.txt (outputs of its own
    `cal_dispersion_curve`, examples 00_0 / 01_0 / 02_0): same rows, same mode labels, c within the float32 accuracy
    the files were generated with (up to 2.6e-6 km/s on config 1; FRIUL7W reaches 4 km/s: 1.2e-6 relative)."""
    from adsurf_b200._utils import gen_model
    from adsurf_b200.workload import friul7w
    fx = load("shipped_fixtures")[name]
    if name == "disper_00":
        mp = api.Model_param(dc=0.001, vmin=2, vmax=3.2, tmin=1 / 10, tmax=5)
        model, num, order = gen_model(np.array([0.4, 0.5, 0.5, 0.5, 0.5, 0.5]),
                                      np.array([2.39, 2.54, 2.7, 2.82, 2.93, 3.04]), area=True), 500, [0]
    elif name == "disper_01":
        mp = api.Model_param(dc=0.001, vmin=0.2, vmax=0.4, tmin=1 / 100, tmax=1 / 5)
        model, num, order = gen_model(np.array([0.005] * 3), np.array([0.3, 0.22, 0.4]), area=True), 500, [0, 1, 2]
    else:
        mp = api.Model_param(dc=0.01, vmin=0.5, vmax=5.0, tmin=1 / 20, tmax=5)
        model, num, order = friul7w(20), 300, [0]
    th, vp, vs, rho = model
    vel = {"thick": th, "vp": vp, "vs": vs, "rho": rho}
    got = api.cal_dispersion_curve(vel, mp, "log-wavelength", num, order, device="Cuda")
    assert got.shape == fx.shape and np.array_equal(got[:, 2], fx[:, 2])
    np.testing.assert_allclose(got[:, 0], fx[:, 0], rtol=1e-6)
    assert np.max(np.abs(got[:, 1] - fx[:, 1])) <= tol
    host = api.cal_dispersion_curve(vel, mp, "log-wavelength", num if name != "disper_01" else 40, order)
    if name != "disper_01":  # the sequential host solver walks the same iterates
        assert np.max(np.abs(host[:, 1] - got[:, 1])) <= 1e-9


def test_dense_grid_extractor_converges_to_roots(api):
    """the other extractor (dense sweep + batched Illinois): zeros of the secular function to rounding, within the
    reference solver's own stopping tolerance (1e-6 relative) of its roots"""
    from adsurf_b200 import ops
    from adsurf_b200.dispersion import dispersion_curves_dense
    g = load("surf96_roots")
    t, d, a, b, rho = (g[f"m6_{k}"] for k in ("t", "d", "a", "b", "rho"))
    c = dispersion_curves_dense(t, d, a, b, rho, modes=(0,), dc=0.001)
    ref = g["m6_mode0_c"]
    assert np.all(c[0] > 0) and np.max(np.abs(c[0] - ref) / ref) < 2e-6
    F = ops.det_points(c[0], t, d, a, b, rho).cpu().numpy()
    cmax, cmin = ops.det_grid_minmax(np.linspace(c[0].min() * 0.9, b.max(), 200), t, d, a, b, rho)
    assert np.max(np.abs(F) / (cmax - cmin).cpu().numpy()) < 1e-9


@pytest.mark.parametrize("case", ["brocher_thick_damped", "constant_thick", "brocher_vs_raw", "usarray_vs_log",
                                  "brocher_thick_LN_nonorm", "usarray_thick_hdamped", "usarray_crossing_4p6"])
def test_fused_optimiser_loop_matches_torch_loop(api, case, monkeypatch):
    """adsurf_ensemble_adam_step (velocity chain + Tikhonov + projection + Adam on the device) against the
    torch-autograd / torch.optim.Adam loop, same order of operations as the reference (_model.py:589-636)."""
    import types
    from adsurf_b200._model import multi_inversion
    from tests.golden_models import brocher_np
    rng = np.random.default_rng(8)
    S, iters = 12, 25
    if case == "constant_thick":
        g = load("dmf_const_m3")
        b = np.tile(g["b"][0], (S, 1)) * (1 + rng.normal(0, 0.03, (S, 3)))
        a, rho, d = b * 2.45, np.full_like(b, 2.0), np.tile(g["d"][0], (S, 1))
        clist, init_method, inv_method, lm = np.arange(0.2, 0.4, 0.002), "Constant", "vs-and-thick", "LN"
        lo, hi, dv, dh, compress, lr = np.full(3, 0.0005), np.full(3, 0.02), 0.0, 0.0, True, 0.001
    else:
        g = load("dmf_exp_m6")
        b = np.tile(g["b"][0], (S, 1)) * (1 + rng.normal(0, 0.03, (S, 6)))
        a, rho = brocher_np(b)
        d = np.tile(g["d"][0], (S, 1))
        clist, init_method = np.arange(2, 3.2, 0.01), "USArray" if case.startswith("usarray") else "Brocher"
        lo, hi = np.full(6, 0.05), np.full(6, 4.0)
        method, normalized = "exp", True
        if case == "brocher_thick_damped":
            inv_method, lm, dv, dh, compress, lr = "vs-and-thick", "LR", 0.002, 0.001, True, 0.005
            lo, hi = np.cumsum(d[0]) * 0.8, np.cumsum(d[0]) * 1.2
        elif case == "usarray_vs_log":
            inv_method, lm, dv, dh, compress, lr, method = "vs", "None", 0.001, 0.0, True, 0.01, "log"
        elif case == "brocher_thick_LN_nonorm":
            inv_method, lm, dv, dh, compress, lr, normalized = "vs-and-thick", "LN", 0.0, 0.0, True, 0.005, False
            lo, hi = d[0] * 0.7, d[0] * 1.4
        elif case == "usarray_thick_hdamped":
            inv_method, lm, dv, dh, compress, lr = "vs-and-thick", "LR", 0.0, 0.002, True, 0.005
            lo, hi = np.cumsum(d[0]) * 0.8, np.cumsum(d[0]) * 1.2
        elif case == "usarray_crossing_4p6":
            # the two deepest layers start either side of 4.6 km/s and move across it: entries above keep the Vp / rho
            # they last had
            b[:, 4] = 4.6 + rng.uniform(-0.03, 0.03, S)
            b[:, 5] = 4.62 + rng.uniform(-0.04, 0.04, S)
            a, rho = brocher_np(b)
            inv_method, lm, dv, dh, compress, lr = "vs", "None", 0.0, 0.0, True, 0.01
        else:
            inv_method, lm, dv, dh, compress, lr = "vs", "None", 0.0, 0.0, False, 0.03
    obs = np.stack([np.column_stack((g["t"][0], g["c"][0]))] * S)
    mp_ = types.SimpleNamespace(clist=clist, layering_method=lm, initialize_method=init_method, vp_vs_ratio=2.45, rho=2)
    if case == "constant_thick":
        method, normalized = "exp", True
    ip = types.SimpleNamespace(damp_vertical=dv, damp_horizontal=dh, compress=compress, normalized=normalized,
                               compress_method=method, inversion_method=inv_method, lr=lr, iteration=iters,
                               step_size=10, gamma=0.5, optimizer="Adam")
    init = types.SimpleNamespace(init_model={"vs": b, "vp": a, "rho": rho, "thick": d, "layer_mindepth": lo,
                                             "layer_maxdepth": hi})
    runs = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("ADSURF_FUSED_OPT", flag)
        ak135 = np.array([[0.0, 2.7, 5.8, 3.4], [35.0, 3.3, 8.0, 4.5], [120.0, 3.4, 8.1, 4.5]])  # depth, rho, vp, vs
        inv = multi_inversion(mp_, ip, init, obs, vsrange_sign="mul", vsrange=[0.7, 1.3], AK135_data=ak135, device="cuda")
        runs[flag] = (np.asarray(inv.inv_process["loss"]), np.asarray(inv.inv_process["iter_vs"]),
                      np.asarray(inv.inv_process["iter_thick"]), np.asarray(inv.inv_model["vs"]))
    # both loops see bit-identical data gradients at identical models; their chain-rule / Adam arithmetic differs in
    # the last bit, and the secular function turns 1 ulp of a thickness into up to ~1e-9 of a gradient
    # (tests/test_host_math.py measures that), so 25 iterations agree to ~1e-8, not to rounding
    for x, y in zip(runs["1"], runs["0"]):
        np.testing.assert_allclose(x, y, rtol=1e-7, atol=1e-10)
    assert np.isfinite(runs["1"][0]).all()
    if case == "usarray_crossing_4p6":
        it = runs["1"][1]  # [iterations, S, L]
        above = it[:, :, 4:] > 4.6
        assert (above[0] != above[-1]).any(), "no entry crossed 4.6 km/s: the case does not test what it says"


@pytest.mark.parametrize("inv_method", ["vs-and-thick", "vs"])
def test_fused_single_model_loop_matches_torch_loop(api, inv_method, monkeypatch):
    """iter_inversion: device-resident iteration vs the torch-autograd loop (same early-stopping bookkeeping)."""
    import types
    from adsurf_b200._model import iter_inversion
    from adsurf_b200._utils import depth_ln
    s = load("shipped_fixtures")
    obs = s["disper_00"][::5]
    lo, hi = depth_ln(0.2, 16.0, 6, 3.8)
    mp_ = types.SimpleNamespace(clist=np.arange(2, 3.2, 0.004), layering_method="LN", initialize_method="Brocher",
                                vp_vs_ratio=2.45, rho=2)
    ip = types.SimpleNamespace(damp_vertical=0.002, damp_horizontal=0, compress=True, normalized=True,
                               compress_method="exp", inversion_method=inv_method, lr=0.005, iteration=30, step_size=10,
                               gamma=0.75, optimizer="Adam")
    init = types.SimpleNamespace(init_model={"vs": s["c2_init_vs"], "vp": s["c2_init_vp"], "rho": s["c2_init_rho"],
                                             "thick": s["c2_init_thick"], "layer_mindepth": lo, "layer_maxdepth": hi})
    runs = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("ADSURF_FUSED_OPT", flag)
        inv = iter_inversion(mp_, ip, init, obs, vsrange_sign="mul", vsrange=[0.1, 2], device="cuda")
        runs[flag] = (np.asarray(inv.inv_process["loss"]), np.asarray(inv.inv_process["iter_vs"]),
                      np.asarray(inv.inv_process["iter_thick"]), np.asarray(inv.inv_model["vs"]))
    for x, y in zip(runs["1"], runs["0"]):
        np.testing.assert_allclose(x, y, rtol=1e-7, atol=1e-10)


def _single_model_problem(iters, lr):
    import types
    from adsurf_b200._utils import depth_ln
    s = load("shipped_fixtures")
    obs = s["disper_00"][::5]
    lo, hi = depth_ln(0.2, 16.0, 6, 3.8)
    mp_ = types.SimpleNamespace(clist=np.arange(2, 3.2, 0.004), layering_method="LN", initialize_method="Brocher",
                                vp_vs_ratio=2.45, rho=2)
    ip = types.SimpleNamespace(damp_vertical=0.002, damp_horizontal=0, compress=True, normalized=True,
                               compress_method="exp", inversion_method="vs-and-thick", lr=lr, iteration=iters,
                               step_size=10, gamma=0.75, optimizer="Adam")
    init = types.SimpleNamespace(init_model={"vs": s["c2_init_vs"], "vp": s["c2_init_vp"], "rho": s["c2_init_rho"],
                                             "thick": s["c2_init_thick"], "layer_mindepth": lo, "layer_maxdepth": hi})
    return mp_, ip, init, obs


def test_early_stopping_of_the_device_resident_loop(api, monkeypatch, capsys):
    """The reference's early stopping (patience 50 on a non-improving loss, _model.py:168-186) evaluated on the host
    over chunks of device-resident iterations stops at the same iteration and returns the same records as the
    per-iteration torch loop (rows past the stopping point stay at their initial values)."""
    from adsurf_b200._model import iter_inversion
    mp_, ip, init, obs = _single_model_problem(iters=90, lr=1e-9)  # the loss barely moves: |dl|/l < 1e-4 every step
    runs = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("ADSURF_FUSED_OPT", flag)
        monkeypatch.setenv("ADSURF_CHECK_EVERY", "7")
        inv = iter_inversion(mp_, ip, init, obs, vsrange_sign="mul", vsrange=[0.1, 2], device="cuda")
        runs[flag] = (np.asarray(inv.inv_process["loss"]), np.asarray(inv.inv_process["iter_vs"]),
                      np.asarray(inv.inv_process["iter_thick"]))
        assert "Early Stopping in Iteration:50" in capsys.readouterr().out
    for x, y in zip(runs["1"], runs["0"]):
        assert x.shape == y.shape
        np.testing.assert_allclose(x, y, rtol=1e-7, atol=1e-10)
    assert np.all(runs["1"][0][51:] == 10.0) and np.all(runs["1"][1][51:] == 0.0)


def test_graph_replay_equals_direct_launches(api):
    """InversionLoop: replaying the captured CUDA graph (unroll 3, 7 iterations = 2 replays + 1 remainder graph)
    gives bit-identical state and history to issuing adsurf_inversion_step call by call; the device iteration
    counter ends at 7; iterating past T leaves the history alone."""
    from adsurf_b200 import _capi
    from tests.golden_models import brocher_np
    rng = np.random.default_rng(11)
    g = load("dmf_exp_m6")
    S, T = 40, 7
    b = np.tile(g["b"][0], (S, 1)) * (1 + rng.normal(0, 0.03, (S, 6)))
    a, rho = brocher_np(b)
    d = np.tile(g["d"][0], (S, 1))
    Gt = lambda x: torch.tensor(np.ascontiguousarray(x, dtype=np.float64), device="cuda")
    t, c = Gt(np.tile(g["t"][0], (S, 1))), Gt(np.tile(g["c"][0], (S, 1)))
    C = Gt(np.tile(np.arange(2, 3.2, 0.01), (S, 1)))
    out = {}
    for use_graph in (False, True):
        state = {"b": Gt(b), "d": Gt(d), "alpha": Gt(a), "rho": Gt(rho)}
        consts = {"b0": Gt(b), "alpha0": Gt(a), "rho0": Gt(rho), "lower_b": Gt(0.7 * b), "upper_b": Gt(1.3 * b),
                  "mindepth": Gt(np.cumsum(d, 1) * 0.8), "maxdepth": Gt(np.cumsum(d, 1) * 1.2)}
        loop = _capi.InversionLoop(state, consts, t, c, C, compress=_capi.COMPRESS_EXP, vp_mode=1, vp_vs_ratio=0.0,
                                   invert_thick=True, layering=1, proj_flags=_capi.PROJ_LAST_LAYER, damp_vertical=0.002,
                                   damp_horizontal=0.001, lr=0.005, gamma=0.5, step_size=3, T=T, unroll=3,
                                   use_graph=use_graph)
        loop.run(T)
        torch.cuda.synchronize()
        assert loop.counter.tolist() == [T, 0]
        hist = (loop.loss_hist.clone(), loop.iter_b_hist.clone(), loop.iter_d_hist.clone())
        loop.run(2)  # past T: state advances, history rows are not written out of bounds
        torch.cuda.synchronize()
        assert loop.counter.tolist() == [T + 2, 0]
        assert all(torch.equal(x, y) for x, y in zip(hist, (loop.loss_hist, loop.iter_b_hist, loop.iter_d_hist)))
        out[use_graph] = hist + (state["b"].clone(), state["d"].clone(), state["alpha"].clone(), state["mb"].clone())
    for x, y in zip(out[True], out[False]):
        assert torch.equal(x, y)
    assert torch.isfinite(out[True][0]).all() and (out[True][0] < 10.0).all()


def test_concurrent_row_groups_equal_one_loop(api):
    """InversionLoopGroup (3 row groups of a 41-model ensemble on 3 streams, CUDA-graph replay each) gives the state
    and the histories of ONE InversionLoop over all rows: the models are independent without horizontal damping."""
    from adsurf_b200 import _capi
    from tests.golden_models import brocher_np
    rng = np.random.default_rng(5)
    g = load("dmf_exp_m6")
    S, T = 41, 7
    b = np.tile(g["b"][0], (S, 1)) * (1 + rng.normal(0, 0.03, (S, 6)))
    a, rho = brocher_np(b)
    d = np.tile(g["d"][0], (S, 1))
    Gt = lambda x: torch.tensor(np.ascontiguousarray(x, dtype=np.float64), device="cuda")
    t, c = Gt(np.tile(g["t"][0], (S, 1))), Gt(np.tile(g["c"][0], (S, 1)))
    C = Gt(np.tile(np.arange(2, 3.2, 0.01), (S, 1)))
    kw = dict(compress=_capi.COMPRESS_EXP, vp_mode=1, vp_vs_ratio=0.0, invert_thick=True, layering=1,
              proj_flags=_capi.PROJ_LAST_LAYER, damp_vertical=0.002, damp_horizontal=0.0, lr=0.005, gamma=0.5,
              step_size=3, T=T, unroll=3)
    out = {}
    for groups in (1, 3):
        state = {"b": Gt(b), "d": Gt(d), "alpha": Gt(a), "rho": Gt(rho)}
        consts = {"b0": Gt(b), "alpha0": Gt(a), "rho0": Gt(rho), "lower_b": Gt(0.7 * b), "upper_b": Gt(1.3 * b),
                  "mindepth": Gt(np.cumsum(d, 1) * 0.8), "maxdepth": Gt(np.cumsum(d, 1) * 1.2)}
        loop = (_capi.InversionLoop(state, consts, t, c, C, **kw) if groups == 1 else
                _capi.InversionLoopGroup(state, consts, t, c, C, groups=groups, **kw))
        loop.run(T)
        torch.cuda.synchronize()
        assert loop.counter.tolist() == [T, 0]
        assert tuple(loop.loss_hist.shape) == (T, S) and tuple(loop.iter_b_hist.shape) == (T, S, 6)
        out[groups] = (loop.loss_hist.clone(), loop.iter_b_hist.clone(), loop.iter_d_hist.clone(), state["b"].clone(),
                       state["d"].clone(), state["alpha"].clone(), state["mb"].clone(), state["vd"].clone())
    for x, y in zip(out[1], out[3]):
        assert torch.equal(x, y)
    with pytest.raises(_capi.AdsurfError):
        _capi.InversionLoopGroup(state, consts, t, c, C, groups=2, **dict(kw, damp_horizontal=0.001))


def test_config4_device_loop_matches_torch_loop_at_1024_models(api, monkeypatch):
    """1024-model FRIUL7W ensemble (N=300 picks, 20 layers, compress=False): every loss, iterate and final model of
    the device-resident loop against the torch-autograd loop over 4 iterations."""
    from adsurf_b200.workload import config5_ensemble
    s = load("shipped_fixtures")
    obs = s["disper_02"]
    d, a, b, rho = config5_ensemble(1024, 20, seed=0)
    mp = api.Model_param(dc=0.01, vmin=0.5, vmax=5.0, tmin=1 / 20, tmax=5, layering_method="None",
                         initialize_method="Brocher", layering_ratio=1.2, depth_factor=2, layer_number=6,
                         vp_vs_ratio=2.45, rho=2)
    ip = api.inv_param(inversion_method="vs", compress=False, compress_method="exp", normalized=True, lr=0.03,
                       damp_vertical=0.0, damp_horizontal=0.0, iteration=4, step_size=2, gamma=0.75)
    init = api.Init_model(model_param=mp, pvs_obs=obs, vp=a[0], vs=b[0], rho=rho[0], thick=d[0])
    for key, val in (("vs", b), ("vp", a), ("thick", d), ("rho", rho)):
        init.init_model[key] = val
    pvs = np.ones((1024, obs.shape[0], 2)) * obs[:, :2]
    runs = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("ADSURF_FUSED_OPT", flag)
        inv = api.inversion(model_param=mp, inv_param=ip, init_model=init, pvs_obs=pvs, vsrange_sign="mul",
                            vsrange=[0.1, 2.5], device="cuda:0")
        assert inv.timing["fused"] == (flag == "1")
        runs[flag] = (np.asarray(inv.inv_process["loss"]), np.asarray(inv.inv_process["iter_vs"]),
                      np.asarray(inv.inv_model["vs"]), np.asarray(inv.inv_model["vp"]))
    for x, y in zip(runs["1"], runs["0"]):
        np.testing.assert_allclose(x, y, rtol=1e-7, atol=1e-10)


def test_nan_propagates_through_dense_and_dmf_kernels(api):
    """NaN inputs never trap: the affected model's outputs are NaN, its neighbours' are untouched (the reference's
    host loop repairs NaN models, _model.py:624-626,638-639)."""
    from adsurf_b200 import _capi
    g = load("grid_m6")
    Gt = lambda x: torch.tensor(np.ascontiguousarray(x, dtype=np.float64), device="cuda")
    d, a, b, rho, t, c = (Gt(g[k]) for k in ("d", "a", "b", "rho", "tlist", "vlist"))
    bad = b.clone()
    bad[1, 2] = float("nan")
    det, cmax, cmin, _ = _capi.grid_fwd(d, a, bad, rho, t, c, want_minmax=True)
    ref = _capi.grid_fwd(d, a, b, rho, t, c)[0]
    assert torch.isnan(det[1]).all() and torch.isnan(cmax[1]).all() and torch.isnan(cmin[1]).all()
    assert torch.equal(det[0], ref[0]) and torch.equal(det[2:], ref[2:])
    out = _capi.grid_bwd(d, a, bad, rho, t, c, None)[1:]
    good = _capi.grid_bwd(d, a, b, rho, t, c, None)[1:]
    for x, y in zip(out, good):
        assert torch.isnan(x[1]).any() and torch.equal(x[0], y[0]) and torch.equal(x[2:], y[2:])
    p = load("dmf_exp_m6")
    pd, pa, pb, pr, pt, pc, pC = (Gt(p[k]) for k in ("d", "a", "b", "rho", "t", "c", "C"))
    pbad = pb.clone()
    pbad[0, 1] = float("nan")
    res = _capi.dmf_step(pd, pa, pbad, pr, pt, pc, pC, _capi.COMPRESS_EXP)
    ok = _capi.dmf_step(pd, pa, pb, pr, pt, pc, pC, _capi.COMPRESS_EXP)
    assert torch.isnan(res[0][0]) and torch.equal(res[0][1:], ok[0][1:])
    for x, y in zip(res[3:], ok[3:]):
        assert torch.equal(x[1:], y[1:])


def test_all_sensitivity_kernels(api):
    """dc/d(vs, vp, thickness, rho) of the 30-layer model from the per-sample Jacobian kernel against autograd on
    the oracle (implicit-function formula of notebook 02_2 cell 12)."""
    from adsurf_b200 import ops
    g = load("points_m30")
    out = ops.sensitivity(g["c"][0], g["t"][0], g["d"][0], g["a"][0], g["b"][0], g["rho"][0])
    Jc = np.diag(g["Jc"]).reshape(-1, 1)
    for key, J in (("dc_dvs", "Jb"), ("dc_dvp", "Ja"), ("dc_dthick", "Jd"), ("dc_drho", "Jrho")):
        ref = -g[J] / Jc
        got = out[key].cpu().numpy()
        scale = np.max(np.abs(ref), axis=1, keepdims=True)
        assert np.max(np.abs(got - ref) / np.maximum(scale, 1e-300)) < 1e-8, key


def test_explicit_init_and_planner_follow_the_device(api):
    from adsurf_b200 import _capi
    lib = _capi.lib()
    assert lib.adsurf_init(0, 0) == 0 and lib.adsurf_init(0, _capi.INIT_PERSIST_L2) == 0
    assert lib.adsurf_init(99, 0) == -6 and lib.adsurf_init(0, 8) == -6
    nsm = torch.cuda.get_device_properties(0).multi_processor_count
    L = 20
    slab = nsm * 12 * (L - 1) * 9 * 32 * 8   # 3 blocks x 4 warps per SM, [L-1][9][32] doubles each
    assert lib.adsurf_workspace_bytes(_capi.OP_GRID_BWD, 4096, L, 128, 1024) >= slab
    assert lib.adsurf_workspace_bytes(_capi.OP_GRID_BWD, 4096, L, 128, 1024) < slab + (64 << 20)


def test_device_side_monte_carlo_ensemble(api):
    """adsurf_mc_generate behind ADsurf_MonteCarlo(generator="device"): same distribution as the reference's numpy
    generator (not the same stream), row 0 = base model, bounds respected, depth rules (LR / LN), deterministic in
    the seed, pre-screen misfit equal to the host-generated path on the same models, and the CUDA-tensor ensemble
    feeds `inversion(...)` without touching the host."""
    from adsurf_b200._MonteCarlo import ADsurf_MonteCarlo
    from adsurf_b200.workload import friul7w
    s = load("shipped_fixtures")
    obs = s["disper_02"]
    th, vp, vs, rho = friul7w(20)
    clist = np.arange(0.5, 5.0, 0.01)
    S = 4096
    kw = dict(layering_method="None", vsrange_sign="add", vsrange=[-1, 1], sampling_num=S, sampling_method="normal",
              gen_velMethod="Brocher", sigma_vs=30, sigma_thick=20)
    res = ADsurf_MonteCarlo(obs, clist, th, vp, vs, rho, forward=True, generator="device", seed=7, **kw)
    M = res["model"]
    assert all(torch.is_tensor(M[k]) and M[k].is_cuda and M[k].shape == (S, 20) for k in ("thick", "vs", "vp", "rho"))
    assert res["loss"].is_cuda and res["loss"].shape == (S,) and torch.isfinite(res["loss"]).all()
    b = M["vs"].cpu().numpy()
    np.testing.assert_array_equal(b[0], vs)
    np.testing.assert_array_equal(M["thick"].cpu().numpy(), np.tile(th, (S, 1)))     # thickness unchanged here
    assert np.all(b >= vs - 1 - 1e-15) and np.all(b <= vs + 1 + 1e-15)
    z = (b[1:] - vs) / (2.0 / 30)                                                      # ~N(0,1): 4095 x 20 draws
    assert abs(z.mean()) < 0.02 and abs(z.std() - 1.0) < 0.02
    assert abs(np.mean(np.abs(z) < 1.0) - 0.6827) < 0.01 and abs(np.mean(z ** 3)) < 0.05
    corr = np.corrcoef(z.T)                                                            # layers are independent
    assert np.max(np.abs(corr - np.eye(20))) < 0.08
    from tests.golden_models import brocher_np
    a_ref, r_ref = brocher_np(b[1:])
    np.testing.assert_allclose(M["vp"].cpu().numpy()[1:], a_ref, rtol=1e-13)
    np.testing.assert_allclose(M["rho"].cpu().numpy()[1:], r_ref, rtol=1e-13)
    again = ADsurf_MonteCarlo(obs, clist, th, vp, vs, rho, forward=False, generator="device", seed=7, **kw)
    other = ADsurf_MonteCarlo(obs, clist, th, vp, vs, rho, forward=False, generator="device", seed=8, **kw)
    assert torch.equal(again["model"]["vs"], M["vs"]) and not torch.equal(other["model"]["vs"], M["vs"])
    # pre-screen on device-generated models == the host path's pre-screen of the same models (first 64)
    from adsurf_b200 import ops
    t = np.tile(obs[:, 0], (64, 1)); c = np.tile(obs[:, 1], (64, 1)); C = np.tile(clist, (64, 1))
    ref = ops.dmf_data_misfit(c, t, C, *(M[k][:64].cpu().numpy() for k in ("thick", "vp", "vs", "rho")), compress="norm")
    np.testing.assert_allclose(res["loss"][:64].cpu().numpy(), ref.cpu().numpy(), rtol=1e-12)
    # depth perturbation: LN keeps every layer at least depth_lower below the previous one; LR clips per layer
    L = 6
    d6, b6 = np.full(L, 0.5), np.linspace(2.4, 3.0, L)
    a6, r6 = brocher_np(b6)
    lo, hi = np.full(L, 0.1), np.cumsum(d6) * 1.5
    for method in ("LR", "LN"):
        out = ADsurf_MonteCarlo(obs, clist, d6, a6, b6, r6, layering_method=method, depth_up_boundary=list(hi),
                                depth_down_boundary=list(lo), vsrange_sign="mul", vsrange=[0.8, 1.2], sampling_num=512,
                                sigma_vs=10, sigma_thick=5, forward=False, generator="device", seed=1)
        thick = out["model"]["thick"].cpu().numpy()
        np.testing.assert_array_equal(thick[0, :-1], d6[:-1])
        assert np.all(thick >= lo[1] - 1e-15) and np.array_equal(thick[:, -1], thick[:, -2])
        assert thick[1:, :-1].std() > 0.01
    # the CUDA-tensor ensemble goes straight into inversion(...)
    mp = api.Model_param(dc=0.01, vmin=0.5, vmax=5.0, tmin=1 / 20, tmax=5, layering_method="None",
                         initialize_method="Brocher", layering_ratio=1.2, depth_factor=2, layer_number=6,
                         vp_vs_ratio=2.45, rho=2)
    ip = api.inv_param(inversion_method="vs", compress=False, lr=0.03, iteration=3, step_size=200, gamma=0.75)
    init = api.Init_model(model_param=mp, pvs_obs=obs, vp=vp, vs=vs, rho=rho, thick=th)
    mc = api.init_model_MonteCarlo(mp, init, sampling_num=256, vsrange_sign="add", vsrange=[-1, 1], sigma_vs=30,
                                   sigma_thick=20, forward=False, generator="device", seed=3)
    for key in ("vs", "vp", "thick", "rho"):
        init.init_model[key] = mc.MonteCarlo_model[key]
    pvs = np.ones((256, obs.shape[0], 2)) * obs[:, :2]
    inv = api.inversion(mp, ip, init, pvs, vsrange_sign="mul", vsrange=[0.1, 2.5], device="cuda")
    assert np.asarray(inv.inv_process["loss"]).shape == (3, 256) and np.isfinite(inv.inv_process["loss"]).all()



def test_inversion_trajectories_match_the_reference(api):
    """loss / iterate trajectories of BASELINE configs 2-4 against the reference's own
    `inversion(...)` run in float64 on the CPU (tests/golden/make_golden_trajectories.py -> trajectories.npz):
    eight iterations through the DMF, the velocity chain, the projections, Adam and StepLR."""
    from tests.trajectories import trajectory_errors
    errs = trajectory_errors(api, "cuda")
    print("trajectory errors vs the reference (float64):", {k: f"{v:.2e}" for k, v in errs.items()})
    tol = 1e-9   # north-star tolerance; measured on B200: <= 4e-12 (losses), <= 4e-13 (iterates)
    bad = {k: v for k, v in errs.items() if not v <= tol}
    assert not bad, f"trajectories differ from the reference's: {bad} (all: {errs})"
