#!/usr/bin/env python
"""End-to-end iterations/s of BASELINE configs 2-4 through the reference-facing API
(`adsurf_b200._ADsurf.inversion`), next to the numbers frozen in the reference's notebooks
(BASELINE.md §1).  Secondary benchmark; the headline metric is bench.py.

    python benchmarks/configs_e2e.py [--iters2 2000] [--iters3 1000] [--iters4 400]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import adsurf_b200._ADsurf as A  # noqa: E402
from adsurf_b200.workload import config5_ensemble  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "shipped_fixtures.npz")


def timed(fn):
    # the reference-style results are nested lists of millions of floats: their collection (or a cyclic-GC pass over
    # them) must not land inside the next call's timers
    import gc
    gc.collect()
    gc.disable()
    sync = torch.cuda.synchronize if torch.cuda.is_available() else (lambda: None)
    try:
        sync()
        t0 = time.perf_counter()
        out = fn()
        sync()
        return out, time.perf_counter() - t0
    finally:
        gc.enable()


def config2(iters):
    s = np.load(GOLD)
    obs = s["disper_00"]
    mp = A.Model_param(dc=0.001, vmin=2, vmax=3.2, tmin=1 / 10, tmax=5, layering_method="LN", initialize_method="Brocher",
                       layering_ratio=1.2, depth_factor=3.8, layer_number=6, vp_vs_ratio=2.45, rho=2)
    ip = A.inv_param(inversion_method="vs-and-thick", compress=True, compress_method="exp", normalized=True, lr=0.005,
                     damp_vertical=0.002, damp_horizontal=0, iteration=iters, step_size=400, gamma=0.75)
    init = A.Init_model(model_param=mp, pvs_obs=obs, thick=[0.4, 0.5, 0.5, 0.5, 0.5, 0.5])
    inv, dt = timed(lambda: A.inversion(mp, ip, init, obs, vsrange_sign="mul", vsrange=[0.1, 2], device="cuda"))
    loss = np.asarray(inv.inv_process["loss"])
    done = int(np.sum(loss < 10)) if len(loss) == iters else len(loss)
    return {"config": 2, "what": "6-layer single model, N=500 picks, K=1200 trial c, vs-and-thick, Adam",
            "iterations_run": len(loss), "it_per_s": len(loss) / dt, "seconds": dt, "best_loss": float(np.nanmin(loss)),
            "first_losses": [float(x) for x in loss[:5]],
            "best_iter": int(np.nanargmin(loss)), "inv_vs": inv.inv_model["vs"],
            "reference": {"it_per_s": 2.36, "device": "cpu (author's)", "best_loss": 1.8169e-4, "best_iter": 1534,
                          "iterations_run": 1722, "true_vs": [2.39, 2.54, 2.7, 2.82, 2.93, 3.04]}}


def config3(iters):
    s = np.load(GOLD)
    obs = s["disper_01_obs"]
    mp = A.Model_param(dc=0.001, vmin=0.2, vmax=0.4, tmin=1 / 100, tmax=1 / 5, layering_method="LN",
                       initialize_method="Constant", layering_ratio=2.5, depth_factor=2.5, layer_number=3, vp_vs_ratio=2.45,
                       rho=2, fundamental_range=[1 / 37, 1 / 8])
    ip = A.inv_param(inversion_method="vs-and-thick", compress=True, compress_method="exp", normalized=True, lr=0.001,
                     damp_vertical=0.0, damp_horizontal=0.0, iteration=iters, step_size=200, gamma=0.75)
    init = A.Init_model(model_param=mp, pvs_obs=obs, thick=np.array([0.005, 0.005, 0.005]))
    init.init_model["thick"][0] -= 0.001
    init.init_model["thick"][1] -= 0.001
    np.random.seed(0)
    mc, dt_mc = timed(lambda: A.init_model_MonteCarlo(mp, init, sampling_num=100, sampling_method="normal",
                                                      vsrange_sign="plus", vsrange=[-0.1, 0.2], sigma_vs=20, sigma_thick=20))
    for key in ("vs", "vp", "thick", "rho"):
        init.init_model[key] = mc.MonteCarlo_model[key]
    pvs = np.ones((100, obs.shape[0], 2)) * obs[:, :2]
    inv, dt = timed(lambda: A.inversion(mp, ip, init, pvs, vsrange_sign="plus", vsrange=[-0.8, 0.8], device="cuda"))
    loss = np.asarray(inv.inv_process["loss"])
    best = np.unravel_index(np.argmin(loss), loss.shape)
    return {"config": 3, "what": "100 models, 3-layer sandwich, N=174 unlabelled multimodal picks, K=200, vs-and-thick",
            "iterations_run": iters, "it_per_s": iters / inv.timing["loop_s"], "loop_seconds": inv.timing["loop_s"],
            "total_seconds_incl_result_lists": dt, "prescreen_models_per_s": 100 / dt_mc,
            "best_loss": float(loss.min()), "best_model_vs": np.asarray(inv.inv_model["vs"])[best[1]].tolist(),
            "first_losses": loss[:5, :4].tolist(),
            "best_model_thick": np.asarray(inv.inv_model["thick"])[best[1]].tolist(),
            "reference": {"it_per_s": 2.20, "device": "cuda (author's)", "best_loss": 2.3655e-4,
                          "prescreen_models_per_s": 23.29, "true_vs": [0.3, 0.22, 0.4], "true_thick": [0.005, 0.005, 0.005]}}


def config4(iters, S=1024, arrays_too=True):
    s = np.load(GOLD)
    obs = s["disper_02"]
    d, a, b, rho = config5_ensemble(S, 20, seed=0)
    mp = A.Model_param(dc=0.01, vmin=0.5, vmax=5.0, tmin=1 / 20, tmax=5, layering_method="None", initialize_method="Brocher",
                       layering_ratio=1.2, depth_factor=2, layer_number=6, vp_vs_ratio=2.45, rho=2)
    ip = A.inv_param(inversion_method="vs", compress=False, compress_method="exp", normalized=True, lr=0.03,
                     damp_vertical=0.0, damp_horizontal=0.0, iteration=iters, step_size=200, gamma=0.75)
    init = A.Init_model(model_param=mp, pvs_obs=obs, vp=a[0], vs=b[0], rho=rho[0], thick=d[0])
    for key, val in (("vs", b), ("vp", a), ("thick", d), ("rho", rho)):
        init.init_model[key] = val
    pvs = np.ones((S, obs.shape[0], 2)) * obs[:, :2]
    inv, dt = timed(lambda: A.inversion(mp, ip, init, pvs, vsrange_sign="mul", vsrange=[0.1, 2.5], device="cuda"))
    loss = np.asarray(inv.inv_process["loss"])
    ip.result_arrays = True  # results as numpy arrays instead of the reference's nested lists
    inv_a, dt_arrays = (timed(lambda: A.inversion(mp, ip, init, pvs, vsrange_sign="mul", vsrange=[0.1, 2.5],
                                                  device="cuda")) if arrays_too else (inv, dt))
    assert np.array_equal(np.asarray(inv_a.inv_process["loss"]), loss)
    return {"config": 4, "total_seconds_result_arrays": dt_arrays, "results_seconds": inv.timing["results_s"],
            "results_seconds_arrays": inv_a.timing["results_s"], "setup_seconds_arrays": inv_a.timing["setup_s"],
            "loop_seconds_arrays": inv_a.timing["loop_s"], "what": f"{S} models, 20-layer FRIUL7W, N=300 picks, compress=False, vs only",
            "iterations_run": iters, "it_per_s": iters / inv.timing["loop_s"], "loop_seconds": inv.timing["loop_s"],
            "gather_seconds": inv.timing["gather_s"], "setup_seconds": inv.timing["setup_s"],
            "total_seconds_incl_result_lists": dt, "best_loss": float(loss.min()),
            "median_first_loss": float(np.median(loss[0])), "median_last_loss": float(np.median(loss[-1])),
            "first_losses": loss[:5, :4].tolist(),
            "kernels_per_iteration": int(getattr(getattr(inv, "loop", None), "kernels_per_iteration", 0)),
            "row_groups": len(getattr(getattr(inv, "loop", None), "loops", [0])),
            "reference": {"it_per_s": 2.60, "models": 100, "device": "cuda:0 (author's)", "best_loss": 2.1835e-4,
                          "iterations": 4000}}


def api_dense_e2e(models=512, steps=3):
    """Config-5 slice through the call a user of the reference makes: the `_cps` shim
    `dltar_matrix(vlist, tlist, d, a, b, rho, 2, -1, device="cuda")` on HOST tensors, `.backward(gdet)`, gradients
    read back to the host.  (The autograd contract costs a forward launch that writes det and a fused
    forward+adjoint launch in backward; inputs go host->device and gradients device->host every step.)"""
    from adsurf_b200._cps import _surf96_matrixAll_gpu as shim
    from adsurf_b200.workload import config5
    d, a, b, rho, t, c = (torch.from_numpy(np.ascontiguousarray(x)) for x in config5(models, 128, 1024, 20, seed=0))
    gdet = torch.ones((models, 1024, 128), dtype=torch.float64, device="cuda")

    def step():
        leaves = [x.clone().requires_grad_(True) for x in (d, a, b, rho)]
        det = shim.dltar_matrix(c, t, *leaves, 2, -1, device="cuda")
        det.backward(gdet)
        return [x.grad.cpu() for x in leaves]

    step()
    _, dt = timed(lambda: [step() for _ in range(steps)])
    n = models * 128 * 1024
    return {"value": n * steps / dt, "unit": "samples/s", "models": models, "steps": steps, "ms_per_step": 1e3 * dt / steps,
            "call": "adsurf_b200._cps._surf96_matrixAll_gpu.dltar_matrix(host tensors, device='cuda').backward(gdet); "
                    "gradients .cpu()",
            "h2d_bytes_per_step": int(sum(x.numel() for x in (d, a, b, rho, t, c)) * 8),
            "d2h_bytes_per_step": int(4 * d.numel() * 8)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters2", type=int, default=2000)
    ap.add_argument("--iters3", type=int, default=1000)
    ap.add_argument("--iters4", type=int, default=400)
    ap.add_argument("--only", type=int, default=0, help="run one config only (2, 3 or 4)")
    ap.add_argument("--models4", type=int, default=1024, help="ensemble size of config 4")
    ap.add_argument("--no-warm", action="store_true", help="skip the warm-up run (profiling: predictable launch counts)")
    args = ap.parse_args()
    A.cal_determinant  # noqa: B018 (import check)
    torch.zeros(1, device="cuda")
    runs = ((2, config2, args.iters2), (3, config3, args.iters3),
            (4, lambda n: config4(n, S=args.models4), args.iters4))
    for k, fn, n in runs:
        if args.only and k != args.only:
            continue
        if not args.no_warm:
            fn(min(n, 5))  # warm-up (library load, allocator)
        print(json.dumps(fn(n)))


if __name__ == "__main__":
    main()
