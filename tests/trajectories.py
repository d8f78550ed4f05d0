"""BASELINE configs 2-4 replayed through the package's `inversion(...)` against the trajectories of the
reference's own `inversion(...)` (float64, CPU: tests/golden/make_golden_trajectories.py -> trajectories.npz).
Shared by the GPU test (real kernels) and the CPU test (C ABI served by the oracle)."""
import numpy as np

from tests.helpers import load


def _traj_err(new, ref):
    """max |new - ref| / max(|ref|, 1e-3 max|ref|): relative, with a floor for entries that pass through zero"""
    new, ref = np.asarray(new, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    assert new.shape == ref.shape, (new.shape, ref.shape)
    return float(np.max(np.abs(new - ref) / np.maximum(np.abs(ref), 1e-3 * np.max(np.abs(ref)))))


def trajectory_errors(api, device, configs=(2, 3, 4), c3_models=100):
    """{name: error} of losses and iterates over the fixture's iterations (8) for the requested configs"""
    g, s = load("trajectories"), load("shipped_fixtures")
    n = g["c2_loss"].shape[0]
    errs = {}
    if 2 in configs:  # single model, vs-and-thick, exp compression, vertical damping, LN depth clipping
        obs = s["disper_00"]
        mp = api.Model_param(dc=0.001, vmin=2, vmax=3.2, tmin=1 / 10, tmax=5, layering_method="LN",
                             initialize_method="Brocher", layering_ratio=1.2, depth_factor=3.8, layer_number=6,
                             vp_vs_ratio=2.45, rho=2)
        ip = api.inv_param(inversion_method="vs-and-thick", compress=True, compress_method="exp", normalized=True,
                           lr=0.005, damp_vertical=0.002, damp_horizontal=0, iteration=n, step_size=400, gamma=0.75)
        init = api.Init_model(model_param=mp, pvs_obs=obs, thick=[0.4, 0.5, 0.5, 0.5, 0.5, 0.5])
        inv = api.inversion(model_param=mp, inv_param=ip, init_model=init, pvs_obs=obs, vsrange_sign="mul",
                            vsrange=[0.1, 2], AK135_data=[], device=device)
        errs.update(c2_loss=_traj_err(inv.inv_process["loss"], g["c2_loss"]),
                    c2_vs=_traj_err(inv.inv_process["iter_vs"], g["c2_iter_vs"]),
                    c2_thick=_traj_err(inv.inv_process["iter_thick"], g["c2_iter_thick"]))
    if 3 in configs:  # 100 models, 3 layers, 174 unlabelled picks, K = 200, vs-and-thick, constant Vp/Vs
        S = c3_models
        obs = s["disper_01_obs"]
        mp = api.Model_param(dc=0.001, vmin=0.2, vmax=0.4, tmin=1 / 100, tmax=1 / 5, layering_method="LN",
                             initialize_method="Constant", layering_ratio=2.5, depth_factor=2.5, layer_number=3,
                             vp_vs_ratio=2.45, rho=2, fundamental_range=[1 / 37, 1 / 8])
        ip = api.inv_param(inversion_method="vs-and-thick", compress=True, compress_method="exp", normalized=True,
                           lr=0.001, damp_vertical=0.0, damp_horizontal=0.0, iteration=n, step_size=200, gamma=0.75)
        init = api.Init_model(model_param=mp, pvs_obs=obs, thick=np.array([0.005, 0.005, 0.005]))
        # the reference offsets the FIRST model's Vs for every model ("plus" bounds): keep model 0 in any subset
        for key in ("vs", "vp", "thick", "rho"):
            init.init_model[key] = g["c3_" + key][:S].copy()
        pvs = np.ones((S, obs.shape[0], 2)) * obs[:, :2]
        inv = api.inversion(model_param=mp, inv_param=ip, init_model=init, pvs_obs=pvs, vsrange_sign="plus",
                            vsrange=[-0.8, 0.8], AK135_data=[], device=device)
        errs.update(c3_loss=_traj_err(inv.inv_process["loss"], g["c3_loss"][:, :S]),
                    c3_vs=_traj_err(inv.inv_process["iter_vs"], g["c3_iter_vs"][:, :S]),
                    c3_thick=_traj_err(inv.inv_process["iter_thick"], g["c3_iter_thick"][:, :S]))
    if 4 in configs:  # FRIUL7W 20-layer ensemble (first 16 of the 1024 models), 300 picks, raw determinants, Vs only
        obs = s["disper_02"]
        d, a, b, rho = g["c4_d"], g["c4_a"], g["c4_b"], g["c4_rho"]
        mp = api.Model_param(dc=0.01, vmin=0.5, vmax=5.0, tmin=1 / 20, tmax=5, layering_method="None",
                             initialize_method="Brocher", layering_ratio=1.2, depth_factor=2, layer_number=6,
                             vp_vs_ratio=2.45, rho=2)
        ip = api.inv_param(inversion_method="vs", compress=False, compress_method="exp", normalized=True, lr=0.03,
                           damp_vertical=0.0, damp_horizontal=0.0, iteration=n, step_size=200, gamma=0.75)
        init = api.Init_model(model_param=mp, pvs_obs=obs, vp=a[0], vs=b[0], rho=rho[0], thick=d[0])
        for key, val in (("vs", b), ("vp", a), ("thick", d), ("rho", rho)):
            init.init_model[key] = val.copy()
        pvs = np.ones((b.shape[0], obs.shape[0], 2)) * obs[:, :2]
        inv = api.inversion(model_param=mp, inv_param=ip, init_model=init, pvs_obs=pvs, vsrange_sign="mul",
                            vsrange=[0.1, 2.5], device=device)
        # model 0 is the true model: its misfit (4.7e-5) is a sum of 300 residuals at the roots, each ~1e-12 of the
        # determinant's scale away from the reference's — compared on the scale of the ensemble's misfits
        loss = np.asarray(inv.inv_process["loss"], dtype=np.float64)
        errs.update(c4_loss=float(np.max(np.abs(loss - g["c4_loss"]) / np.maximum(np.abs(g["c4_loss"]), 1.0))),
                    c4_vs=_traj_err(inv.inv_process["iter_vs"], g["c4_iter_vs"]))
    return errs
